// K3/K4: fused attention core on tcgen05 (see attention_tcgen05.cu).
#pragma once
#include "runtime.h"

namespace tsd {
bool attention_fused_supported(int d, int causal);
int attention_fused(Ctx* c, const AttnArgs& a);
}  // namespace tsd
