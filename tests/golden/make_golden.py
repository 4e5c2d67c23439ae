"""Generate the golden fixtures under tests/golden/ (run in the build container only).

Three kinds of pins:
 1. ``ref_index_maps.npz`` -- integer maps produced by executing the reference's OWN pure-NumPy
    lines verbatim (text read from /root/reference/modules.py at generation time, never copied
    into this repo): ``relative_position_index`` (modules.py:88-98) and the shifted-window region
    image (modules.py:192-203) followed by a NumPy transliteration of window_partition
    (modules.py:49-55; the only TF ops there are reshape/transpose) and the mask rule (:209-212).
 2. ``oracle_outputs.npz`` -- subsampled outputs of oracle/strajnet_oracle.py on seeded inputs, so a
    later edit of the oracle that changes its numbers is caught.  (The reference itself cannot
    run here: no TensorFlow.  See the oracle header, "parity unpinned".)
 3. ``eval_outputs.npz`` -- the same drift guard for oracle/eval_oracle.py (validation-side loss / metrics);
    ``--eval-only`` regenerates just this file.

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/modules.py"


def ref_lines(lo, hi):
    with open(REF) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[lo - 1:hi]))


def ref_relative_position_index(ws):
    class S:  # stands in for `self`
        window_size = (ws, ws)
    env = {"np": np, "self": S}
    exec(ref_lines(88, 98), env)
    return env["relative_position_index"]


def ref_shift_mask(H, W, ws, shift):
    class S:
        input_resolution = (H, W)
        window_size = ws
        shift_size = shift
    env = {"np": np, "self": S}
    exec(ref_lines(191, 203), env)
    img = env["img_mask"]  # [1,H,W,1] float64 region ids
    mw = img.reshape(1, H // ws, ws, W // ws, ws, 1).transpose(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    diff = mw[:, None, :] - mw[:, :, None]
    return np.where(diff != 0, -100.0, 0.0)


def main():
    out = {}
    rpi = ref_relative_position_index(8)
    out["relative_position_index_ws8"] = rpi
    print("rpi sha256", hashlib.sha256(np.ascontiguousarray(rpi).tobytes()).hexdigest()[:16], rpi.sum())
    for H in (16, 32, 64, 128):
        m = ref_shift_mask(H, H, 8, 4)
        print(H, "mask sha256", hashlib.sha256(m.astype(np.float32).tobytes()).hexdigest()[:16], int((m != 0).sum()))
        out[f"shift_mask_{H}"] = m.astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "ref_index_maps.npz"), **out)

    from oracle import strajnet_oracle as O
    torch.set_num_threads(8)
    gold = {}
    for shift in (0, 4):
        for heads in (1, 2):
            w = O.make_block_weights(32, heads, seed=0)
            rng = np.random.Generator(np.random.PCG64(0))
            x = torch.from_numpy(rng.standard_normal((1, 4096, 32)).astype(np.float32))
            y = O.swin_block(x, w, "", 64, 64, heads, 8, shift)
            gold[f"block_c32_h{heads}_s{shift}"] = y[0, ::37].numpy()
    w = O.make_weights(O.CFG256, seed=0)
    inp = O.make_inputs(1, 256, seed=0)
    y = O.forward_from_inputs(w, O.CFG256, inp)
    gold["forward_cfg256_fg"] = y[0, ::16, ::16].numpy()
    y = O.forward_from_inputs(w, O.CFG256, inp, fg_msa=False, fg=False)
    gold["forward_cfg256_nofg"] = y[0, ::16, ::16].numpy()
    np.savez_compressed(os.path.join(HERE, "oracle_outputs.npz"), **gold)
    print({k: v.shape for k, v in gold.items()})


EVAL_FLAGS = {"default": dict(), "train_py": dict(use_gt=True, use_focal_loss=False), "use_pred": dict(use_pred=True)}


def eval_golden():
    """``eval_outputs.npz`` -- drift guard for oracle/eval_oracle.py (row f4): the four losses for three constructor
    configurations and the seven metrics on the seeded synthetic batch (B = 2, 64 x 64)."""
    from oracle import eval_oracle as E
    d = E.make_eval_inputs(2, 64, seed=0)
    gold = {}
    for name, flags in EVAL_FLAGS.items():
        r = E.ogm_flow_loss(**d, **flags)
        gold[f"loss_{name}"] = np.array([r[k].item() for k in ("observed_xe", "occluded_xe", "flow", "flow_warp_xe")], np.float64)
    m = E.occupancy_flow_metrics(**d)
    gold["metrics"] = np.array([v.item() for v in m.values()], np.float64)
    np.savez_compressed(os.path.join(HERE, "eval_outputs.npz"), **gold)
    print({k: v.tolist() for k, v in gold.items()})


if __name__ == "__main__":
    import sys
    if "--eval-only" in sys.argv:
        eval_golden()
    else:
        main()
        eval_golden()
