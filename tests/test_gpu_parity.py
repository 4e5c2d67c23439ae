"""GPU parity tests: every layer of the CUDA path (through the C ABI, via the Python mirror of the
reference's Keras classes) against the CPU oracle on the same seeded inputs.

Tolerances: bit-exact for integer / index maps; <= 1e-3 abs for fp32 (north_star); bf16 is reported
and gated loosely (it is not the parity configuration).
"""
import numpy as np
import pytest
import torch

from oracle import strajnet_oracle as O
from tests.util import max_abs, oracle_model, randn, sub

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-3
CFG256 = O.CFG256


@pytest.fixture(scope="module")
def sj():
    import strajnet_b200
    assert torch.cuda.is_available()
    return strajnet_b200


# ------------------------------------------------------------------------------------ integer maps
def test_relative_position_index_bit_exact(sj, golden_dir):
    g = np.load(f"{golden_dir}/ref_index_maps.npz")
    out = sj.relative_position_index(8).cpu().numpy()
    assert out.dtype == np.int64
    assert np.array_equal(out, g["relative_position_index_ws8"])


@pytest.mark.parametrize("H", [16, 32, 64, 128])
def test_shift_mask_bit_exact(sj, golden_dir, H):
    g = np.load(f"{golden_dir}/ref_index_maps.npz")
    out = sj.shift_attn_mask(H, H, 8, 4).cpu().numpy()
    assert np.array_equal(out, g[f"shift_mask_{H}"].astype(np.float32))


@pytest.mark.parametrize("H,W,shift", [(16, 16, 0), (16, 16, 4), (64, 64, 4), (32, 64, 4), (128, 128, 0)])
def test_window_token_map_bit_exact(sj, H, W, shift):
    idx = torch.arange(H * W, dtype=torch.float32).reshape(1, H, W, 1)
    if shift:
        idx = torch.roll(idx, (-shift, -shift), (1, 2))
    ref = O.window_partition(idx, 8).reshape(-1).to(torch.int32)
    out = sj.window_token_map(H, W, 8, shift).cpu()
    assert torch.equal(out, ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_window_partition_reverse_bit_exact(sj, dtype):
    x = randn((3, 16, 24, 32), 1).to(dtype)
    w = sj.window_partition(x.cuda(), 8)
    assert torch.equal(w.cpu(), O.window_partition(x, 8))
    back = sj.window_reverse(w, 8, 16, 24, 32)
    assert torch.equal(back.cpu(), x)


# ------------------------------------------------------------------------------------ Swin pieces
def test_mlp(sj):
    w = O.make_block_weights(96, 3, seed=5)
    m = sj.Mlp(96, 384)
    m.set_weights(sub(w, "mlp."))
    x = randn((2, 100, 96), 6)
    ref = O.dense(O.gelu_tanh(O.dense(x, w["mlp.fc1.kernel"], w["mlp.fc1.bias"])), w["mlp.fc2.kernel"], w["mlp.fc2.bias"])
    assert max_abs(m(x), ref) < FP32_TOL


@pytest.mark.parametrize("C,heads,masked", [(32, 2, False), (32, 1, True), (96, 3, True), (384, 12, False)])
def test_window_attention(sj, C, heads, masked):
    w = O.make_block_weights(C, heads, seed=7)
    layer = sj.WindowAttention(C, (8, 8), heads)
    layer.set_weights(sub(w, "attn."))
    nW = 4
    x = randn((2 * nW, 64, C), 8)
    mask = torch.from_numpy(O.shift_attn_mask(16, 16, 8, 4)).float() if masked else None
    ref = O.window_attention(x, w, "attn.", heads, 8, mask)
    assert max_abs(layer(x, mask=mask), ref) < FP32_TOL
    assert torch.equal(layer.relative_position_index.cpu(), torch.from_numpy(O.relative_position_index(8)))


@pytest.mark.parametrize("heads", [1, 2])
@pytest.mark.parametrize("shift", [0, 4])
def test_swin_block_config1(sj, golden_dir, heads, shift):
    """BASELINE config 1: SwinTransformerBlock(dim=32, (64,64), window 8), batch 1."""
    w = O.make_block_weights(32, heads, seed=0)
    blk = sj.SwinTransformerBlock(32, (64, 64), heads, window_size=8, shift_size=shift)
    blk.set_weights(w)
    x = randn((1, 4096, 32), 0)
    y = blk(x)
    ref = O.swin_block(x, w, "", 64, 64, heads, 8, shift)
    assert max_abs(y, ref) < FP32_TOL
    g = np.load(f"{golden_dir}/oracle_outputs.npz")[f"block_c32_h{heads}_s{shift}"]
    assert np.abs(y[0, ::37].cpu().numpy() - g).max() < FP32_TOL
    if shift:
        assert torch.equal(blk.attn_mask.cpu(), torch.from_numpy(O.shift_attn_mask(64, 64, 8, 4)).float())


@pytest.mark.parametrize("C,heads,H,B", [(96, 3, 64, 2), (192, 6, 32, 2), (384, 12, 16, 3), (96, 3, 8, 1)])
@pytest.mark.parametrize("shift", [0, 4])
def test_swin_block_model_shapes(sj, C, heads, H, B, shift):
    w = O.make_block_weights(C, heads, seed=11)
    blk = sj.SwinTransformerBlock(C, (H, H), heads, window_size=8, shift_size=shift)
    blk.set_weights(w)
    x = randn((B, H * H, C), 12)
    ref = O.swin_block(x, w, "", H, H, heads, 8, shift)
    assert max_abs(blk(x), ref) < FP32_TOL


def test_swin_block_rejects_wrong_length(sj):
    blk = sj.SwinTransformerBlock(32, (64, 64), 2, window_size=8)
    with pytest.raises(AssertionError):
        blk(randn((1, 4000, 32)))


def test_patch_merging(sj):
    w = oracle_model()
    p = "encoder.basic_layers.0.downsample."
    layer = sj.PatchMerging((64, 64), 96)
    layer.set_weights(sub(w, p))
    x = randn((2, 4096, 96), 13)
    assert max_abs(layer(x), O.patch_merging(x, w, p, 64, 64)) < FP32_TOL


@pytest.mark.parametrize("name,cin", [("vecicle", 11), ("map", 3), ("flow", 2)])
def test_patch_embed(sj, name, cin):
    w = oracle_model()
    p = f"encoder.patch_embed_{name}."
    layer = sj.PatchEmbed((256, 256), (4, 4), cin, 96)
    layer.set_weights(sub(w, p))
    x = randn((2, 256, 256, cin), 14)
    assert max_abs(layer(x), O.patch_embed(x, w, p)) < FP32_TOL


def test_basic_layer(sj):
    w = oracle_model()
    p = "encoder.basic_layers.1."
    layer = sj.BasicLayer(192, (32, 32), 2, 6, 8, downsample=True)
    layer.set_weights(sub(w, p))
    x = randn((2, 1024, 192), 15)
    yd, res = layer(x)
    rd, rres = O.basic_layer(x, w, p, 32, 32, 2, 6, 8, True)
    assert max_abs(res, rres) < FP32_TOL and max_abs(yd, rd) < FP32_TOL


@pytest.mark.parametrize("S,large", [(256, False), (512, True)])
def test_encoder(sj, S, large):
    cfg = O.CFG512 if large else O.CFG256
    w = oracle_model()
    enc = sj.SwinTransformerEncoder(img_size=cfg["input_size"], window_size=8, embed_dim=96, depths=[2, 2, 2],
                                    num_heads=[3, 6, 12], sep_encode=True, flow_sep=True, use_flow=True, large_input=large)
    enc.set_weights(sub(w, "encoder."))
    inp = O.make_inputs(1, S, seed=2)
    outs = enc(inp["ogm"], inp["map_img"], inp["flow"], training=False)
    refs = O.encoder_forward(inp["ogm"], inp["map_img"], inp["flow"], w, cfg, large)
    for o, r in zip(outs, refs):
        assert o.numel() == r.numel()  # res2 is [B,16,16,384] here; the 512 oracle path re-flattens it
        assert max_abs(o.reshape(r.shape), r) < FP32_TOL


# ------------------------------------------------------------------------------------ FG-MSA, trajectories, decoder
def test_fgmsa(sj):
    w = oracle_model()
    layer = sj.FGMSA((16, 16), (16, 16), 8, 48, n_groups=8, out_dim=384, fg=True)
    layer.set_weights(sub(w, "fg_msa_layer."))
    x = randn((2, 16, 16, 384), 16)
    y, pos, hid = layer(x, training=False)
    ry, rpos, rhid = O.fgmsa_forward(x, w)
    assert max_abs(pos, rpos) < FP32_TOL
    assert max_abs(hid, rhid) < FP32_TOL
    assert max_abs(y, ry) < FP32_TOL


def _traj_inputs(B, seed, mode="normal"):
    inp = O.make_inputs(B, 256, seed=seed)
    obs, occ = inp["obs"], inp["occ"]
    if mode == "no_occ":
        occ = torch.zeros_like(occ)
    elif mode == "all_padded":  # every actor padded: all attention rows fully masked -> uniform (Q7)
        obs, occ = torch.zeros_like(obs), torch.zeros_like(occ)
    return obs, occ


@pytest.mark.parametrize("mode", ["normal", "no_occ", "all_padded"])
def test_traj_cross_attention(sj, mode):
    w = oracle_model()
    layer = sj.TrajNetCrossAttention(dict(traj_heads=4, att_heads=6, out_dim=384, no_attn=False), pic_size=(16, 16), pic_dim=384)
    layer.set_weights(sub(w, "trajnet_attn."))
    B = 2
    pic = randn((B, 8, 16, 16, 384), 17)
    obs, occ = _traj_inputs(B, 3, mode)
    out = layer(pic, obs, occ, None, training=False)
    ref = O.trajnet_cross_attention(pic, obs, occ, w)
    assert max_abs(out, ref) < FP32_TOL


def test_decoder(sj):
    w = oracle_model()
    dec = sj.Pyramid3DDecoder(None, (256, 256), use_pyramid=True, timestep_split=True, shallow_decode=1,
                              flow_sep_decode=True, conv_cnn=False)
    dec.set_weights(sub(w, "decoder."))
    B = 1
    x = randn((B, 8, 16, 16, 384), 18)
    res = [randn((B, 4096, 96), 19), randn((B, 4096, 96), 20), randn((B, 1024, 192), 21), randn((B, 16, 16, 384), 22)]
    out = dec(x, training=False, res_list=res)
    ref = O.decoder_forward(x, res, w)
    assert tuple(out.shape) == (B, 8, 256, 256, 4)
    assert max_abs(out, ref) < FP32_TOL


# ------------------------------------------------------------------------------------ whole forward
def _model(sj, fg_msa=True, fg=True, large=False, dtype="float32"):
    cfg = O.CFG512 if large else O.CFG256
    m = sj.STrajNet(cfg, fg_msa=fg_msa, fg=fg, large_ogm=large, dtype=dtype)
    m.set_weights(oracle_model(fg_msa=fg_msa, fg=fg))
    return m


def _fwd(m, inp):
    return m(inp["ogm"], inp["map_img"], training=False, obs=inp["obs"], occ=inp["occ"], mapt=inp["mapt"], flow=inp["flow"])


def test_strajnet_config2_fp32(sj, golden_dir):
    """BASELINE config 2: full forward, 256x256, 8 waypoints, batch 1, fp32, <= 1e-3 abs vs the oracle."""
    m = _model(sj)
    inp = O.make_inputs(1, 256, seed=0)
    y = _fwd(m, inp)
    assert tuple(y.shape) == (1, 256, 256, 32) and y.dtype == torch.float32
    ref = O.forward_from_inputs(oracle_model(), CFG256, inp)
    err = max_abs(y, ref)
    print(f"config2 fp32 max abs err vs oracle: {err:.3e}")
    assert err < FP32_TOL
    g = np.load(f"{golden_dir}/oracle_outputs.npz")["forward_cfg256_fg"]
    assert np.abs(y[0, ::16, ::16].cpu().numpy() - g).max() < FP32_TOL


def test_strajnet_without_fgmsa(sj):
    m = _model(sj, fg_msa=False, fg=False)
    inp = O.make_inputs(1, 256, seed=4)
    ref = O.forward_from_inputs(oracle_model(fg_msa=False, fg=False), CFG256, inp, fg_msa=False, fg=False)
    assert max_abs(_fwd(m, inp), ref) < FP32_TOL


def test_strajnet_batch_invariance_and_parity_b3(sj):
    m = _model(sj)
    inp = O.make_inputs(3, 256, seed=5)
    y = _fwd(m, inp)
    ref = O.forward_from_inputs(oracle_model(), CFG256, inp)
    assert max_abs(y, ref) < FP32_TOL
    one = {k: v[1:2] for k, v in inp.items()}
    assert torch.equal(_fwd(m, one)[0], y[1])  # sample i of a batch == the batch-1 result, bit for bit


def test_strajnet_config5_large_input(sj):
    """BASELINE config 5 geometry (512x512 input, large_ogm=True) at batch 1."""
    m = _model(sj, large=True)
    inp = O.make_inputs(1, 512, seed=6)
    ref = O.forward_from_inputs(oracle_model(), O.CFG512, inp, large_ogm=True)
    assert max_abs(_fwd(m, inp), ref) < FP32_TOL


def test_strajnet_bf16_reported(sj):
    """BASELINE config 3 arithmetic (bf16 activations, fp32 accumulate): error reported, gated loosely."""
    m = _model(sj, dtype="bfloat16")
    inp = O.make_inputs(2, 256, seed=7)
    y = _fwd(m, inp)
    ref = O.forward_from_inputs(oracle_model(), CFG256, inp)
    err = max_abs(y, ref)
    rel = err / ref.abs().max().item()
    print(f"bf16 max abs err vs fp32 oracle: {err:.3e} (rel to max |y|: {rel:.3e})")
    assert torch.isfinite(y).all() and rel < 0.08


def test_empty_and_bad_inputs(sj):
    m = _model(sj)
    inp = O.make_inputs(1, 256, seed=0)
    with pytest.raises(ValueError):
        m(inp["ogm"][:, :128], inp["map_img"], training=False, obs=inp["obs"], occ=inp["occ"], flow=inp["flow"])
    with pytest.raises(ValueError):
        m(inp["ogm"], inp["map_img"], training=False, obs=inp["obs"][:, :10], occ=inp["occ"], flow=inp["flow"])
    with pytest.raises(NotImplementedError):
        m(inp["ogm"], inp["map_img"], obs=inp["obs"], occ=inp["occ"], flow=inp["flow"])  # training defaults to True
    z = {k: torch.zeros_like(v) for k, v in inp.items()}  # the reference's own dummy-zero build call
    ref = O.forward_from_inputs(oracle_model(), CFG256, z)
    assert max_abs(_fwd(m, z), ref) < FP32_TOL


def test_inference_pipeline_matches_direct_call(sj):
    """The pipelined serving API returns the same logits as a direct model call, for every batch in flight."""
    from strajnet_b200.pipeline import InferencePipeline
    m = _model(sj)
    pipe = InferencePipeline(m, batch=2)
    batches = [O.make_inputs(2, 256, seed=20 + i) for i in range(5)]
    outs = [h.result().clone() for h in pipe.run({k: v.pin_memory() for k, v in b.items() if k != "mapt"} for b in batches)]
    assert len(outs) == 5
    for b, y in zip(batches, outs):
        assert torch.equal(y, _fwd(m, b).cpu())


def _raw_inputs(B, seed):
    """The record's own types (inference.py:91-93): ogm bool bytes, map int8; plus their float decode."""
    inp = O.make_inputs(B, 256, seed=seed)
    ogm_u8 = (inp["ogm"] != 0).to(torch.uint8)
    map_i8 = torch.round(inp["map_img"] * 256).to(torch.int8)
    ogm_f, map_f = O.decode_raw_inputs(ogm_u8, map_i8)
    assert torch.equal(ogm_f, inp["ogm"]) and torch.equal(map_f, inp["map_img"])
    return inp, ogm_u8, map_i8


def test_raw_typed_inputs_bit_identical(sj):
    """f3: feeding bool/uint8 rasters and the int8 map directly equals feeding their float decode, bit for bit."""
    m = _model(sj)
    inp, ogm_u8, map_i8 = _raw_inputs(2, 31)
    y_f = _fwd(m, inp)
    y_r = m(ogm_u8, map_i8, training=False, obs=inp["obs"], occ=inp["occ"], flow=inp["flow"])
    assert torch.equal(y_f, y_r)
    y_b = m(ogm_u8.bool(), map_i8, training=False, obs=inp["obs"], occ=inp["occ"], flow=inp["flow"])
    assert torch.equal(y_f, y_b)


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_fused_submission_quantisation(sj, dtype):
    """f1: the fused uint8/int8 epilogue equals the oracle's quantiser (inference.py:124-136,160-182) applied to the
    same model's fp32 logits; rounding-boundary flips of one LSB are tolerated on < 0.1 % of the bytes."""
    m = _model(sj, dtype=dtype)
    inp, ogm_u8, map_i8 = _raw_inputs(2, 32)
    logits = _fwd(m, inp).cpu()
    q = m.predict_quantized(ogm_u8, map_i8, inp["obs"], inp["occ"], inp["flow"]).cpu()
    ref = O.quantize_outputs(logits)
    assert q.dtype == torch.uint8 and tuple(q.shape) == (2, 256, 256, 32)
    # compare occupancy as uint8 and flow as int8
    qa, ra = q.reshape(-1, 8, 4), ref.reshape(-1, 8, 4)
    d_occ = (qa[..., :2].int() - ra[..., :2].int()).abs()
    d_flw = (qa[..., 2:].view(torch.int8).int() - ra[..., 2:].view(torch.int8).int()).abs()
    assert d_occ.max() <= 1 and d_flw.max() <= 1
    frac = ((d_occ > 0).float().mean().item() + (d_flw > 0).float().mean().item()) / 2
    print(f"{dtype}: fraction of bytes off by one LSB: {frac:.2e}")
    assert frac < 1e-3
    if dtype == "float32":  # and against the oracle's own logits
        ref2 = O.quantize_outputs(O.forward_from_inputs(oracle_model(), CFG256, inp))
        d2 = (q.reshape(-1, 8, 4)[..., :2].int() - ref2.reshape(-1, 8, 4)[..., :2].int()).abs()
        assert d2.max() <= 1


def test_pipeline_raw_quantised(sj):
    from strajnet_b200.pipeline import InferencePipeline
    m = _model(sj)
    pipe = InferencePipeline(m, batch=2, raw_inputs=True, quantized=True)
    inp, ogm_u8, map_i8 = _raw_inputs(2, 33)
    host = dict(ogm=ogm_u8.pin_memory(), map_img=map_i8.pin_memory(), obs=inp["obs"].pin_memory(),
                occ=inp["occ"].pin_memory(), flow=inp["flow"].pin_memory())
    outs = [h.result().clone() for h in pipe.run(host for _ in range(3))]
    ref = m.predict_quantized(ogm_u8, map_i8, inp["obs"], inp["occ"], inp["flow"]).cpu()
    for y in outs:
        assert torch.equal(y, ref)


def test_records_to_model_bit_identical(sj, tmp_path):
    """f3 end to end: TFRecord file -> framing -> Example -> record decode (strajnet_b200/records.py) -> model.  The raw
    record dtypes (bool ogm, int8 map) fed straight to the device give the same logits, bit for bit, as the reference's
    float32 decode (inference.py:84-96); config = the reference's inference config (512 input, inference.py:142)."""
    from strajnet_b200 import records as R
    from tests.test_records import _synthetic_example
    path = str(tmp_path / "scenes.tfrecords")
    R.write_tfrecords(path, [_synthetic_example(s)[1] for s in (11, 12)])
    recs = list(R.read_tfrecords(path))
    raw = R.batch_examples([R.decode_example(r, raw=True) for r in recs])
    ref = R.batch_examples([R.decode_example(r, raw=False) for r in recs])
    assert raw["scenario/id"] == [b"scene11", b"scene12"]
    w = O.make_weights(O.CFG512, seed=0)
    m = sj.STrajNet(O.CFG512, fg_msa=True, fg=True, large_ogm=True, dtype="bfloat16")
    m.set_weights(w)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    y_raw = m(t(raw["ogm"]), t(raw["map_img"]), training=False, obs=t(raw["obs"]), occ=t(raw["occ"]), mapt=t(raw["mapt"]), flow=t(raw["flow"]))
    y_ref = m(t(ref["ogm"]), t(ref["map_img"]), training=False, obs=t(ref["obs"]), occ=t(ref["occ"]), mapt=t(ref["mapt"]), flow=t(ref["flow"]))
    assert y_raw.shape == (2, 256, 256, 32) and torch.isfinite(y_raw).all()
    assert torch.equal(y_raw, y_ref)


def test_strajnet_tf_checkpoint_roundtrip(sj, tmp_path):
    """f2: `save_weights(prefix)` writes a TF-format checkpoint (tensor bundle + object graph), `load_weights(prefix)` on a
    fresh model restores every parameter by attribute path; the two models agree bit for bit."""
    from strajnet_b200 import tf_checkpoint as T
    m = _model(sj)
    prefix = str(tmp_path / "final_model")
    m.save_weights(prefix)
    b = T.TensorBundle(prefix)
    assert len(b.variables_by_attribute_path()) == len(m.get_weights()) == 299
    m2 = sj.STrajNet(O.CFG256, fg_msa=True, fg=True, large_ogm=False)
    m2.load_weights(prefix)
    inp = O.make_inputs(1, 256, seed=41)
    assert torch.equal(_fwd(m, inp), _fwd(m2, inp))


def test_forward_into_cuda_graph_replay(sj):
    """forward_into(graph=True): first use launches + captures, later uses replay; results equal the plain launches
    bit for bit, also after the inputs were overwritten in place and after set_weights invalidated the capture."""
    m = _model(sj, dtype="bfloat16")
    dev = m.device
    inp = {k: v.to(dev) for k, v in O.make_inputs(2, 256, seed=51).items()}
    out = torch.empty(2, 256, 256, 32, device=dev)
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        args = (inp["ogm"], inp["map_img"], inp["obs"], inp["occ"], inp["flow"])
        ref = m.forward_into(torch.empty_like(out), *args).clone()
        for _ in range(3):  # capture, replay, replay
            out.zero_()
            m.forward_into(out, *args, graph=True)
            assert torch.equal(out, ref)
        assert len(m._graphs) == 1
        new = {k: v.to(dev) for k, v in O.make_inputs(2, 256, seed=52).items()}
        for k in inp:
            inp[k].copy_(new[k])  # same buffers, new contents
        ref2 = m.forward_into(torch.empty_like(out), *args).clone()
        m.forward_into(out, *args, graph=True)
        assert torch.equal(out, ref2) and not torch.equal(ref, ref2)
        m.set_weights(O.make_weights(O.CFG256, seed=1))
        assert len(m._graphs) == 0
        ref3 = m.forward_into(torch.empty_like(out), *args).clone()
        m.forward_into(out, *args, graph=True)
        m.forward_into(out, *args, graph=True)
        assert torch.equal(out, ref3)
    s.synchronize()


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_programmatic_dependent_launch_is_invisible(sj, dtype):
    """Every kernel orders itself behind its predecessor with griddepcontrol.wait, so launching the forward with
    programmatic stream serialisation (the default) must give the same bits as fully serialised launches: stream launches
    and graph replay, several runs each (a missing wait would show up as a race on the reused workspace)."""
    from strajnet_b200 import _lib
    lib = _lib.lib()
    m = _model(sj, dtype=dtype)
    dev = m.device
    inp = {k: v.to(dev) for k, v in O.make_inputs(3, 256, seed=61).items()}
    args = (inp["ogm"], inp["map_img"], inp["obs"], inp["occ"], inp["flow"])
    s = torch.cuda.Stream(dev)
    prev = lib.sj_set_pdl(0)
    try:
        with torch.cuda.stream(s):
            ref = m.forward_into(torch.empty(3, 256, 256, 32, device=dev), *args).clone()
            assert lib.sj_set_pdl(3) == 0
            out = torch.empty_like(ref)
            for _ in range(3):
                out.zero_()
                m.forward_into(out, *args)
                assert torch.equal(out, ref)
            for _ in range(3):  # capture with programmatic edges, then replay
                out.zero_()
                m.forward_into(out, *args, graph=True)
                assert torch.equal(out, ref)
        s.synchronize()
    finally:
        lib.sj_set_pdl(prev)


def test_headline_batch16_bf16_batch_invariance(sj):
    """BASELINE config 3 at full size (batch 16, bf16) through a size-independent property: every sample of the batch
    equals the same sample run alone, bit for bit (no cross-sample op anywhere in the path), for the logits
    and for the fused submission quantisation."""
    m = _model(sj, dtype="bfloat16")
    inp = O.make_inputs(16, 256, seed=71)
    y = _fwd(m, inp)
    assert tuple(y.shape) == (16, 256, 256, 32) and torch.isfinite(y).all()
    for i in (0, 7, 15):
        one = {k: v[i:i + 1] for k, v in inp.items()}
        assert torch.equal(_fwd(m, one)[0], y[i]), f"sample {i} differs between batch 16 and batch 1"


@pytest.mark.parametrize("name,B,S,large", [("config3_b2", 2, 256, False), ("config3_b16", 16, 256, False),
                                             ("config5_b4", 4, 512, True)])
def test_strajnet_bf16_error_profile(sj, name, B, S, large):
    """The benchmarked arithmetic (bf16 tensor-core path) against the fp32 oracle at the benchmarked sizes, gated on the
    error DISTRIBUTION rather than on the maximum alone: mean, 99.9th percentile, and the image border against the
    interior (a wrong border tap of a sub-pixel phase or a mis-placed halo shows up as a border / interior gap long before
    it moves the maximum of 2 M logits)."""
    m = _model(sj, dtype="bfloat16", large=large)
    inp = O.make_inputs(B, S, seed=31)
    y = _fwd(m, inp).float().cpu()
    ref = O.forward_from_inputs(oracle_model(), O.CFG512 if large else CFG256, inp, large_ogm=large)
    err = (y - ref).abs()
    rng = ref.abs().max().item()
    flat = err.flatten()
    p999 = flat.kthvalue(int(0.999 * flat.numel())).values.item()
    border = torch.zeros(256, 256, dtype=torch.bool)
    border[0, :] = border[-1, :] = border[:, 0] = border[:, -1] = True
    rms_b = err[:, border].pow(2).mean().sqrt().item()
    rms_i = err[:, ~border].pow(2).mean().sqrt().item()
    print(f"bf16 {name}: max {flat.max().item():.3e} ({flat.max().item() / rng:.3%} of range {rng:.2f}), mean {flat.mean().item():.3e}, "
          f"p99.9 {p999:.3e}, rms border {rms_b:.3e} / interior {rms_i:.3e}")
    assert torch.isfinite(y).all()
    assert flat.max().item() < 0.08 * rng
    assert flat.mean().item() < 2.5e-2 and p999 < 0.1
    assert rms_b < 1.6 * rms_i + 1e-3
    # The noise floor of bf16 storage itself: the same graph evaluated by the oracle with every kernel and every stored
    # activation rounded to bf16 at layer granularity (O.bf16_storage) deviates from the fp32 oracle by as much as the CUDA
    # path does -- the two bf16 evaluations are nearly uncorrelated (rounding noise through ~70 layers, not a systematic
    # offset), so neither is a tighter reference for the other; what CAN be asserted is that the CUDA path is no noisier than
    # an independent bf16 evaluation of the reference graph.  (Bit-level models exist per kernel: tests/test_gpu_kernels.py.)
    with O.bf16_storage():
        emu = O.forward_from_inputs(oracle_model(), O.CFG512 if large else CFG256, inp, large_ogm=large)
    eflat = (emu - ref).abs().flatten()
    e999 = eflat.kthvalue(int(0.999 * eflat.numel())).values.item()
    rms_c, rms_e = flat.pow(2).mean().sqrt().item(), eflat.pow(2).mean().sqrt().item()
    print(f"bf16 {name}: rms {rms_c:.3e} vs {rms_e:.3e} for the bf16-storage oracle (mean {eflat.mean().item():.3e}, p99.9 {e999:.3e}, "
          f"max {eflat.max().item():.3e}); CUDA vs bf16-storage oracle rms {(y - emu).pow(2).mean().sqrt().item():.3e}")
    assert rms_c < 1.3 * rms_e and flat.mean().item() < 1.3 * eflat.mean().item()
    assert p999 < 1.3 * e999 and flat.max().item() < 1.5 * eflat.max().item()


def test_inference_pipeline_stale_handle_raises(sj):
    """A handle whose slot has been handed to a later submit() must not return that batch's data."""
    from strajnet_b200.pipeline import InferencePipeline
    m = _model(sj)
    pipe = InferencePipeline(m, batch=1, depth=2)
    b = {k: v.pin_memory() for k, v in O.make_inputs(1, 256, seed=40).items() if k != "mapt"}
    h0 = pipe.submit(b)
    h1 = pipe.submit(b)
    h2 = pipe.submit(b)  # reuses h0's slot
    with pytest.raises(RuntimeError):
        h0.result()
    assert torch.equal(h1.result(), h2.result())
    pipe.synchronize()


def test_vehicle_plane_only_input_is_bit_identical(sj):
    """ogm handed over as its vehicle plane alone ([B,S,S,11], uint8 or fp32) gives the same bits as the record's
    [B,S,S,11,2] raster, of which the model reads plane 0 only (modules.py:572)."""
    for dtype in ("float32", "bfloat16"):
        m = _model(sj, dtype=dtype)
        inp = O.make_inputs(2, 256, seed=41)
        y = _fwd(m, inp)
        plane = inp["ogm"][..., 0].contiguous()
        y1 = m(plane, inp["map_img"], training=False, obs=inp["obs"], occ=inp["occ"], flow=inp["flow"])
        y2 = m((plane != 0).to(torch.uint8), inp["map_img"], training=False, obs=inp["obs"], occ=inp["occ"], flow=inp["flow"])
        assert torch.equal(y, y1) and torch.equal(y, y2)


def test_bf16_forward_is_run_to_run_deterministic(sj):
    """No kernel of the bf16 path uses atomics or order-dependent reductions, so repeated forwards must agree bit for bit,
    also when the batch size changes in between (new workspace, new tensor maps, different tiles per CTA).  This is the
    guard for hand-off races in the TMA / mbarrier pipelines: the first TMA-fed version of head_tapsum_kernel passed every
    parity gate in isolation and corrupted a few hundred logits per batch-16 forward in this sequence."""
    m = _model(sj, dtype="bfloat16")
    for B in (1, 2, 16, 1, 16):
        inp = O.make_inputs(B, 256, seed=31)
        y0 = _fwd(m, inp).clone()
        for _ in range(3):
            assert torch.equal(_fwd(m, inp), y0), f"batch {B}: two forwards of the same inputs differ"
