"""SURVEY section 8 row f2 (infrastructure): the safetensors reader of the C ABI and the blob assembly, on the CPU.
The reference has no loader and no checkpoint exists offline, so the checks are against numpy: the reader returns
exactly what the writer stored (F32 bit-exact, F16 / BF16 widened exactly), for every malformed file an error."""
import json
import struct

import numpy as np
import pytest

from tsd_b200 import weights as W
from tsd_b200._lib import TsdError


def tensors(seed=0):
    rng = np.random.default_rng(seed)
    t = {"unet.layer1.weight": rng.standard_normal((8, 4, 3, 3)).astype(np.float32),
         "unet.layer1.bias": rng.standard_normal(8).astype(np.float32),
         "time_embed.layer1.weight": rng.standard_normal((16, 5)).astype(np.float32),
         "scalar": np.float32(3.5).reshape(()),
         "empty": np.zeros((0, 4), np.float32),
         "specials": np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 6.1e-5, 5.96e-8, 65504.0, 1e-40, -1e38], np.float32)}
    return t


@pytest.mark.parametrize("dtype", ["F32", "F16", "BF16"])
def test_reader_returns_what_was_written(tmp_path, dtype):
    t = tensors()
    p = tmp_path / "w.safetensors"
    with np.errstate(over="ignore"):                     # 1e38 -> F16 overflows to inf on purpose
        W.write_safetensors(p, t, dtype, {"note": 'quote " and \\ backslash', "n": 3})
    for src in (p, p.read_bytes()):                      # memory-mapped file and in-memory copy
        st = W.SafeTensors(src)
        assert st.names() == list(t)
        for name, a in t.items():
            d, shape = st.info(name)
            assert d == dtype and shape == a.shape
            got = st.read(name)
            if dtype == "F32":
                want = a
            elif dtype == "F16":
                with np.errstate(over="ignore"):
                    want = a.astype(np.float16).astype(np.float32)
            else:
                want = (W._bf16_bits(a).astype(np.uint32) << 16).view(np.float32).reshape(a.shape)
            assert got.dtype == np.float32 and got.shape == a.shape
            bits = lambda v: np.asarray(v, np.float32).reshape(-1).view(np.uint32)  # noqa: E731
            assert np.array_equal(bits(got), bits(want)), (name, dtype)
        st.close()


def test_f64_and_non_float_tensors(tmp_path):
    a = np.array([1.5, -2.25, 1e300], np.float64)
    i = np.arange(6, dtype=np.int32)
    hdr = {"d": {"dtype": "F64", "shape": [3], "data_offsets": [0, 24]},
           "i": {"dtype": "I32", "shape": [2, 3], "data_offsets": [24, 48]}}
    hj = json.dumps(hdr).encode()
    blob = struct.pack("<Q", len(hj)) + hj + a.tobytes() + i.tobytes()
    st = W.SafeTensors(blob)
    with np.errstate(over="ignore"):
        assert np.array_equal(st.read("d"), a.astype(np.float32))
    assert st.info("i") == ("I32", (2, 3))
    with pytest.raises(TsdError):
        st.read("i")                                     # listed, but not a floating-point parameter


def test_malformed_files_are_errors(tmp_path):
    t = {"a": np.arange(6, dtype=np.float32).reshape(2, 3)}
    p = tmp_path / "w.safetensors"
    W.write_safetensors(p, t)
    good = p.read_bytes()
    (hlen,) = struct.unpack_from("<Q", good)
    hdr = json.loads(good[8:8 + hlen])

    def rebuild(h, data=good[8 + hlen:]):
        hj = json.dumps(h).encode()
        return struct.pack("<Q", len(hj)) + hj + data

    bad = [good[:4], b"", struct.pack("<Q", 1 << 40) + good[8:], good[:9] + b"x" + good[10:],
           rebuild({"a": {"dtype": "F32", "shape": [2, 3], "data_offsets": [0, 20]}}),      # size != shape x dtype
           rebuild({"a": {"dtype": "F32", "shape": [2, 3], "data_offsets": [8, 32]}}),      # past the end
           rebuild({"a": {"dtype": "F32", "shape": [2, 3]}}),                               # no offsets
           rebuild({"a": {"dtype": "F32", "shape": [2, -3], "data_offsets": [0, 24]}}),
           rebuild(hdr, good[8 + hlen:-4])]                                                 # truncated data
    for b in bad:
        with pytest.raises(TsdError):
            W.SafeTensors(b)
    with pytest.raises(TsdError):
        W.SafeTensors(tmp_path / "missing.safetensors")
    assert W.SafeTensors(rebuild({})).names() == []
    assert W.SafeTensors(rebuild({"__metadata__": {"k": "v"}})).names() == []


def test_build_blob_from_name_map(tmp_path):
    t = tensors(3)
    p = tmp_path / "ckpt.safetensors"
    W.write_safetensors(p, {"model.diffusion_model.input_blocks.0.0.weight": t["unet.layer1.weight"],
                            "model.diffusion_model.input_blocks.0.0.bias": t["unet.layer1.bias"],
                            "time_embed.layer1.weight": t["time_embed.layer1.weight"],
                            "first_stage_model.unrelated": np.ones(3, np.float32)}, "F16")
    table = [("unet.layer1.weight", 0, 288), ("unet.layer1.bias", 320, 8), ("time_embed.layer1.weight", 384, 80)]
    name_map = {"unet.layer1.weight": "model.diffusion_model.input_blocks.0.0.weight",
                "unet.layer1.bias": "model.diffusion_model.input_blocks.0.0.bias"}
    st = W.SafeTensors(p)
    blob, rep = W.build_blob(table, st, name_map)
    h = lambda a: a.astype(np.float16).astype(np.float32).reshape(-1)  # noqa: E731
    assert blob.shape == (464,) and rep == {"missing": [], "unused": ["first_stage_model.unrelated"]}
    assert np.array_equal(blob[:288], h(t["unet.layer1.weight"])) and np.array_equal(blob[320:328], h(t["unet.layer1.bias"]))
    assert np.array_equal(blob[384:464], h(t["time_embed.layer1.weight"])) and not blob[288:320].any()
    with pytest.raises(TsdError):                        # a parameter the checkpoint does not hold
        W.build_blob(table + [("unet.layer2.weight", 464, 4)], st, name_map)
    _, rep = W.build_blob(table + [("unet.layer2.weight", 464, 4)], st, name_map, strict=False)
    assert rep["missing"] == ["unet.layer2.weight"]
    with pytest.raises(TsdError):                        # element count mismatch
        W.build_blob([("unet.layer1.bias", 0, 9)], st, name_map)


def test_tiny_sd_name_map_round_trip(tmp_path):
    """A synthetic checkpoint under diffusers-convention names (the layout segmind/tiny-sd ships in) assembles, through
    tiny_sd_unet_name_map, into exactly the blob a norm_affine Diffusion expects: every parameter is found, the fused
    self-attention in_proj is rebuilt from to_q | to_k | to_v, conv kernels keep OIHW, and the tensors the reference's
    topology has no home for (upsampler convolutions) are reported as unused instead of being dropped silently."""
    import synth
    specs = synth.diffusion_specs(norm_affine=True)
    blob = synth.random_blob(specs, 21)
    ours = synth.BlobWeights(specs, blob)
    table, off = [], 0
    for name, shape, _ in specs:
        n = int(np.prod(shape))
        table.append((name, off, n))
        off += n
    nm = W.tiny_sd_unet_name_map([t[0] for t in table])
    assert len(nm) == len(table) and len(set(map(str, nm.values()))) == len(nm)     # one-to-one
    assert nm["unet.layer2.layer1.weight"] == "down_blocks.0.resnets.0.norm1.weight"
    assert nm["unet.layer23.layer8.bias"] == "up_blocks.2.attentions.1.transformer_blocks.0.ff.net.0.proj.bias"
    assert nm["unet.layer4.weight"] == "down_blocks.0.downsamplers.0.conv.weight"
    assert nm["final.layer1.bias"] == "conv_norm_out.bias"
    file_tensors = {}
    for name, shape, _ in specs:
        src = nm[name]
        if isinstance(src, list):                       # split the fused projection the way the file stores it
            parts = np.split(ours[name], len(src), axis=0)
            for s_, p_ in zip(src, parts):
                file_tensors[s_] = p_
        elif src.endswith("proj_in.weight") or src.endswith("proj_out.weight"):
            file_tensors[src] = ours[name].reshape(shape[0], shape[1])   # tiny-sd stores the 1x1 projections as linears
        else:
            file_tensors[src] = ours[name]
    file_tensors["up_blocks.0.upsamplers.0.conv.weight"] = np.zeros((4, 4, 3, 3), np.float32)   # no counterpart (Q8)
    path = tmp_path / "tiny_sd_like.safetensors"
    W.write_safetensors(path, file_tensors)
    st = W.SafeTensors(path)
    got, report = W.build_blob(table, st, nm)
    st.close()
    assert np.array_equal(got, blob)
    assert report["missing"] == [] and report["unused"] == ["up_blocks.0.upsamplers.0.conv.weight"]


@pytest.mark.parametrize("which", ["decoder", "encoder", "clip"])
def test_tiny_sd_vae_and_clip_name_maps_round_trip(tmp_path, which):
    """The VAE (diffusers AutoencoderKL names) and the text encoder (transformers CLIPTextModel names) of a
    segmind/tiny-sd style checkpoint assemble into the blobs of TSD_MODEL_NORM_AFFINE models: one-to-one maps, fused
    in_proj rebuilt from q | k | v, nothing missing, nothing unused."""
    import synth
    specs, mapper = {"decoder": (synth.decoder_specs(norm_affine=True), W.tiny_sd_vae_decoder_name_map),
                     "encoder": (synth.encoder_specs(norm_affine=True), W.tiny_sd_vae_encoder_name_map),
                     "clip": (synth.clip_specs(300, 2, norm_affine=True), W.tiny_sd_clip_name_map)}[which]
    blob = synth.random_blob(specs, 23)
    ours = synth.BlobWeights(specs, blob)
    table, off = [], 0
    for name, shape, _ in specs:
        n = int(np.prod(shape))
        table.append((name, off, n))
        off += n
    nm = mapper([t[0] for t in table])
    assert len(nm) == len(table) and len(set(map(str, nm.values()))) == len(nm)
    if which == "decoder":
        assert nm["l4.groupnorm.weight"] == "decoder.mid_block.attentions.0.group_norm.weight"
        assert nm["l16.res_conv_layer.bias"] == "decoder.up_blocks.2.resnets.0.conv_shortcut.bias"
        assert nm["l24.bias"] == "decoder.conv_norm_out.bias" and nm["l1.weight"] == "post_quant_conv.weight"
    elif which == "encoder":
        assert nm["l19.weight"] == "quant_conv.weight" and nm["l10.bias"] == "encoder.down_blocks.2.downsamplers.0.conv.bias"
        assert nm["l14.attention.out_proj.weight"] == "encoder.mid_block.attentions.0.to_out.0.weight"
    else:
        assert nm["player2.layer2.in_proj.bias"] == [f"text_model.encoder.layers.1.self_attn.{p}_proj.bias" for p in "qkv"]
        assert nm["layernorm.weight"] == "text_model.final_layer_norm.weight"
    file_tensors = {}
    for name, shape, _ in specs:
        src = nm[name]
        if isinstance(src, list):
            for s_, p_ in zip(src, np.split(ours[name], len(src), axis=0)):
                file_tensors[s_] = p_
        else:
            file_tensors[src] = ours[name]
    path = tmp_path / f"{which}.safetensors"
    W.write_safetensors(path, file_tensors)
    st = W.SafeTensors(path)
    got, report = W.build_blob(table, st, nm)
    st.close()
    assert np.array_equal(got, blob)
    assert report == {"missing": [], "unused": []}
