#include "attention_tcgen05.cuh"
namespace tsd {
bool attention_fused_supported(int d, int causal) { (void)d; (void)causal; return false; }
int attention_fused(Ctx* c, const AttnArgs& a) { (void)a; return c->fail(TSD_ERR_STATE, "fused attention not built"); }
}  // namespace tsd
