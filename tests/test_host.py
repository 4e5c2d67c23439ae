"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, struct
layouts agree, argument validation works without a GPU, weight preparation is consistent."""
import ctypes as C
import os
import re

import pytest
import torch

from strajnet_b200 import _lib as L
from strajnet_b200 import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "strajnet_b200.h")).read()
    declared = set(re.findall(r"\b(sj_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert lib.sj_version() >= 100


def test_struct_layouts_match():
    lib = L.lib()
    structs = [L.SjLinear, L.SjNorm, L.SjSwinBlockW, L.SjPatchMergeW, L.SjPatchEmbedW, L.SjBasicLayerW, L.SjEncoderW,
               L.SjFgmsaW, L.SjTrajW, L.SjDecoderW, L.SjModelW]
    for i, s in enumerate(structs):
        assert lib.sj_sizeof(i) == C.sizeof(s)


def test_status_strings_and_errors():
    lib = L.lib()
    assert lib.sj_strerror(0) == b"ok"
    for s in (-1, -2, -3, -4):
        assert lib.sj_strerror(s) not in (b"ok", b"unknown status")
    with pytest.raises(ValueError):
        L.check(L.SJ_EINVAL, "x")
    with pytest.raises(L.SjError):
        L.check(L.SJ_EWORKSPACE, "x")
    # null pointers are rejected before anything touches the device
    assert lib.sj_strajnet_fwd(None, None, None, None, None, None, None, 1, 256, 0, None, 0, None) == L.SJ_EINVAL
    assert lib.sj_relative_position_index(8, None, None) == L.SJ_EINVAL


def test_workspace_queries_scale_with_batch():
    lib = L.lib()
    w1 = lib.sj_strajnet_workspace_bytes(1, 256, L.SJ_F32)
    w4 = lib.sj_strajnet_workspace_bytes(4, 256, L.SJ_F32)
    assert 0 < w1 < w4 <= 4 * w1 + (1 << 20)
    assert lib.sj_strajnet_workspace_bytes(4, 256, L.SJ_BF16) < w4
    assert lib.sj_swin_block_workspace_bytes(1, 64, 64, 32, L.SJ_F32) > 4096 * 32 * 4 * 9


def test_packer_builds_model_struct_on_cpu():
    from oracle import strajnet_oracle as O
    w = O.make_weights(O.CFG256, seed=0)
    assert set(W.model_shapes(O.CFG256, True, True)) == set(O.weight_shapes(O.CFG256, True, True))
    p = W.Packer(w, "cpu", tc=True)
    m = p.model(O.CFG256, True, True, False)
    assert m.encoder.num_layers == 3 and m.encoder.layers[2].dim == 384 and m.encoder.layers[2].heads == 12
    assert m.decoder.upconv[0].w and m.decoder.upconv[0].w_tc and m.traj.ca_q.w
    assert m.fg_msa == 1 and m.large_ogm == 0


def test_tfa_relayout():
    k = torch.arange(3 * 5 * 4, dtype=torch.float32).reshape(3, 5, 4)
    m = W.tfa_in_kernel(k, 16)
    assert m.shape == (5, 16) and torch.equal(m[:, 4:8], k[1]) and torch.all(m[:, 12:] == 0)
    x = torch.randn(7, 5)
    assert torch.allclose(x @ m[:, :12], torch.einsum("ni,hio->nho", x, k).reshape(7, 12))
    pk = torch.randn(3, 4, 6)
    o = torch.randn(7, 3, 4)
    assert torch.allclose(o.reshape(7, 12) @ W.tfa_out_kernel(pk), torch.einsum("nhi,hio->no", o, pk), atol=1e-6)


def test_layer_argument_validation():
    import strajnet_b200 as sj
    with pytest.raises(ValueError):
        sj.STrajNet(dict(input_size=(256, 256), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12]),
                    large_ogm=True, device="cpu")
    with pytest.raises(ValueError):
        sj.STrajNet(dict(input_size=(256, 256), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12]),
                    large_ogm=False, fg=True, fg_msa=False, device="cpu")
    m = sj.STrajNet(dict(input_size=(256, 256), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12]),
                    large_ogm=False, fg_msa=True, fg=True, device="cpu")
    with pytest.raises(NotImplementedError):
        m(None, None, training=True)
    with pytest.raises(AssertionError):
        sj.SwinTransformerBlock(32, (64, 64), 2, window_size=8, shift_size=8, device="cpu")
    blk = sj.SwinTransformerBlock(32, (4, 4), 2, window_size=8, shift_size=4, device="cpu")
    assert blk.window_size == 4 and blk.shift_size == 0  # modules.py:173-175


def test_evaluation_host_logic():
    """Row f4 host side (no GPU): the WaypointGrids container packs to the kernel's layouts exactly as train.py:103-140
    slices them, the constructor switches map to the C-ABI flag bits, and CPU tensors are refused (no CPU fallback)."""
    import pytest
    import torch
    from strajnet_b200 import _lib as L
    from strajnet_b200 import evaluation as V
    from strajnet_b200.loss import OGMFlow_loss
    from strajnet_b200.occu_metric import WaypointGrids, compute_occupancy_flow_metrics  # noqa: F401

    g = torch.Generator().manual_seed(0)
    out = torch.randn(2, 8, 8, 32, generator=g)
    gt = dict(obs=torch.rand(2, 8, 8, 8, 1, generator=g), occ=torch.rand(2, 8, 8, 8, 1, generator=g),
              flow=torch.randn(2, 8, 8, 8, 2, generator=g), org=torch.rand(2, 8, 8, 8, 1, generator=g))
    pred, true = WaypointGrids(), WaypointGrids()
    for k in range(8):  # train.py:103-121 and :126-140
        c = out[:, :, :, 4 * k: 4 * k + 4]
        pred.vehicles.observed_occupancy.append(c[:, :, :, :1])
        pred.vehicles.occluded_occupancy.append(c[:, :, :, 1:2])
        pred.vehicles.flow.append(c[:, :, :, 2:])
        true.vehicles.observed_occupancy.append(gt["obs"][:, k])
        true.vehicles.occluded_occupancy.append(gt["occ"][:, k])
        true.vehicles.flow.append(gt["flow"][:, k])
        true.vehicles.flow_origin_occupancy.append(gt["org"][:, k])
    assert torch.equal(V.pack_predictions(pred, device="cpu"), out)
    o, c, f, r = V.pack_truth(true, device="cpu")
    assert torch.equal(o, gt["obs"][..., 0]) and torch.equal(c, gt["occ"][..., 0])
    assert torch.equal(f, gt["flow"]) and torch.equal(r, gt["org"][..., 0])
    pred.vehicles.flow.pop()
    with pytest.raises(ValueError):
        V.pack_predictions(pred, device="cpu")
    # flag bits = the enum of include/strajnet_b200.h
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "strajnet_b200.h")).read()
    for name, bit in (("USE_FOCAL", "use_focal"), ("NO_USE_WARP", "no_use_warp"), ("USE_PRED", "use_pred"), ("USE_GT", "use_gt"),
                      ("PRED_IS_PROB", "pred_is_prob"), ("LOSS", "loss"), ("METRICS", "metrics"),
                      ("METRICS_NO_WARP", "metrics_no_warp")):
        assert f"SJ_EVAL_{name} = {V.FLAG[bit]}" in hdr
    assert f"#define SJ_EVAL_OUT_FLOATS {V.OUT_FLOATS}" in hdr
    assert OGMFlow_loss(None).flags() == V.FLAG["use_focal"]  # reference defaults (loss.py:24-25)
    assert OGMFlow_loss(None, use_gt=True, use_focal_loss=False, no_use_warp=True).flags() == V.FLAG["use_gt"] | V.FLAG["no_use_warp"]
    assert C.sizeof(L.SjEvalParams) == 20
    with pytest.raises(RuntimeError):
        OGMFlow_loss(None).packed(out, o, c, f, r)  # CPU tensors: the op runs on the GPU or raises


def test_output_grid_answers_the_serving_loops_calls():
    """inference.py:109-113 slices the model output, :130 applies an op to the slices, :169-181 calls .numpy() on them."""
    import numpy as np
    import torch
    from strajnet_b200.layers import OutputGrid
    y = torch.arange(2 * 4 * 4 * 32, dtype=torch.float32).reshape(2, 4, 4, 32).as_subclass(OutputGrid)
    k = 3
    wp = y[:, :, :, k * 4:(k + 1) * 4]
    obs, flow = wp[:, :, :, :1], wp[:, :, :, 2:]
    assert isinstance(obs, OutputGrid) and isinstance(torch.sigmoid(obs), OutputGrid)
    a = torch.sigmoid(obs).numpy()
    assert isinstance(a, np.ndarray) and a.shape == (2, 4, 4, 1)
    q = np.clip(np.round(flow.numpy()), -128, 127).astype(np.int8)
    assert q.shape == (2, 4, 4, 2)
    assert np.asarray(wp).dtype == np.float32 and np.asarray(wp, dtype=np.float64).dtype == np.float64


def test_numa_binding_is_best_effort():
    from strajnet_b200.parallel import bind_to_local_numa
    assert isinstance(bind_to_local_numa(0), dict)  # no GPU here: {} and no exception


def test_fg_conv_fragment_order():
    """weights.fg_conv_fragments: element (group, tap, k-step, n-pair, lane, e) of the packed copy is the conv kernel entry
    the mma.m16n8k16 B fragment of fg_offset_mma.cu expects there (FG_MSA.py:51 kernel, [3,3,48,384])."""
    import torch
    from strajnet_b200.weights import fg_conv_fragments
    g = torch.Generator().manual_seed(5)
    k = torch.randn(3, 3, 48, 384, generator=g)
    f = fg_conv_fragments(k).reshape(8, 9, 3, 3, 32, 8)
    w = k.reshape(9, 48, 384)
    idx = torch.randint(0, 10 ** 6, (400, 6), generator=g)
    for row in idx.tolist():
        grp, tap, ks, pr, lane, e = (row[i] % n for i, n in enumerate((8, 9, 3, 3, 32, 8)))
        gq, t = lane // 4, lane % 4
        kk = 16 * ks + 2 * t + (0, 1, 8, 9)[e & 3]
        n = 48 * grp + 8 * (2 * pr + (e >> 2)) + gq
        assert f[grp, tap, ks, pr, lane, e] == w[tap, kk, n]
