"""Microbenchmarks of BASELINE.json configs[3] (RoI-warp + mask-pool, HBM GB/s vs roofline) and
configs[4] (gpu_nms 10k boxes / keep 300 + gpu_mask_voting, bit-exact vs the oracle)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mnc_b200 import ops
from tests import util


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for i in range(iters):
        if flush is not None:
            flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    # ---- config 4: 2000 RoIs on a 512x38x63 map (SURVEY.md section 8d inputs)
    g = torch.Generator(device="cpu").manual_seed(7)
    feat = torch.randn(1, 512, 38, 63, generator=g).clamp_min(0).cuda()
    rng = np.random.default_rng(8)
    x1 = rng.uniform(0, 999, 2000); y1 = rng.uniform(0, 599, 2000)
    w = rng.uniform(16, 600, 2000); h = rng.uniform(16, 600, 2000)
    rois = np.stack([np.zeros(2000), x1, y1, np.clip(x1 + w, 0, 999), np.clip(y1 + h, 0, 599)], 1).astype(np.float32)
    trois = torch.from_numpy(rois).cuda()
    for P in (28, 14):
        out = torch.empty(2000, 512, P, P, device="cuda")
        ms = timeit(lambda: ops.roi_warp_nchw(feat, trois, P, P, out=out), flush=flush)
        alg = 2000 * 512 * P * P * 4 + 512 * 38 * 63 * 4 + 2000 * 20
        res["roi_warp_P%d" % P] = {"ms": ms, "algorithmic_bytes": alg, "GBps": alg / ms / 1e6,
                                   "frac_of_measured_hbm": alg / ms / 1e6 / hbm}
    f14 = torch.randn(2000, 512, 14, 14, device="cuda")
    m14 = torch.rand(2000, 1, 14, 14, device="cuda")
    o14 = torch.empty_like(f14)
    ms = timeit(lambda: ops.mask_pool_nchw(f14, m14, out=o14), flush=flush)
    alg = 2 * 2000 * 512 * 196 * 4 + 2000 * 196 * 4
    res["mask_pool"] = {"ms": ms, "algorithmic_bytes": alg, "GBps": alg / ms / 1e6,
                        "frac_of_measured_hbm": alg / ms / 1e6 / hbm}
    # ---- config 5: NMS 10k boxes / keep 300 (device form, sorted input) + host drop-in
    boxes = util.random_boxes(10000, seed=10)
    scores = util.tie_free_scores(10000, seed=11)
    order = np.argsort(-scores, kind="stable")
    sb = torch.from_numpy(boxes[order]).cuda()[None].contiguous()
    ms = timeit(lambda: ops.nms_sorted(sb, None, 0.7, 300))
    res["nms_10k_keep300_device"] = {"ms": ms, "pair_ious": 10000 * 10000 // 2}
    ms = timeit(lambda: ops.nms_sorted(sb, None, 0.7, 0))
    res["nms_10k_keep_all_device"] = {"ms": ms}
    import mnc_b200.lib as L
    L.install()
    from nms.gpu_nms import gpu_nms
    from oracle import oracle as O
    dets = np.hstack([boxes, scores[:, None]]).astype(np.float32)
    t0 = time.perf_counter(); keep = gpu_nms(dets, 0.7); t1 = time.perf_counter()
    want = O.gpu_nms(dets, 0.7)
    res["nms_10k_host_dropin"] = {"ms": (t1 - t0) * 1e3, "kept": len(keep),
                                  "bit_exact_vs_oracle": [int(k) for k in keep] == [int(k) for k in want]}
    # ---- config 5b: mask voting, 600 boxes x 21 classes at 600x1000
    from tests.test_ref_pin import _voting_inputs
    vb, vm, vs = _voting_inputs(600, 600, 1000, 11)
    tb, tm, ts = (torch.from_numpy(a).cuda()[None] for a in (vb, vm, vs))
    hw = torch.tensor([[600, 1000]], dtype=torch.int32, device="cuda")
    ms = timeit(lambda: ops.mask_voting(tb, tm, ts, hw), iters=5)
    r = ops.mask_voting(tb, tm, ts, hw)
    inds, start, wts, cs, bar = O.mask_voting_candidates(vb, vs, 21, 100)
    k = int(r["n_res"][0])
    beg = r["cand_begin"][0, :k].cpu().numpy(); end = r["cand_end"][0, :k].cpu().numpy()
    ci = r["cand_inds"][0].cpu().numpy().ravel()
    got = np.concatenate([ci[b:e] for b, e in zip(beg, end)])
    res["mask_voting_600x21"] = {"ms": ms, "results": k, "candidates": int(len(inds)),
                                 "lists_bit_exact_vs_oracle": bool(np.array_equal(got, inds))}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
