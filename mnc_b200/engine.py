"""Batched MNC 5-stage inference engine (the hot path of BASELINE.json's north_star).

One `MNCEngine` per GPU: weights resident in HBM as split-bf16, activations NHWC split-bf16,
every layer of models/VGG16/mnc_5stage/test.prototxt executed by a kernel of libmnc_b200.so on
torch's current stream, no host round trips inside `forward` (the reference crosses D<->H four
times per image for its Python layers: SURVEY.md section 3.2).  The reference is batch-1
(lib/pylayer/proposal_layer.py:65); here images are batched by looping the proposal stage per
image on device and stacking RoIs with their batch index (ROIWarping honours roi[0],
roi_warping_layer.cu:79,96), so the FC weights stream from HBM once per batch.

PyTorch's role: device memory and streams (a step is ~60 asynchronous launches queued in 1.2 ms
of host time against 15 ms of device time, so no CUDA graph is needed to keep the GPU busy).  Apart
from a few `torch.cat` / copies in the `detect` tail, no torch op computes on the path.
"""
import math

import torch

from . import dense, ops
from .weights import TRUNK_NAMES, POOL_AFTER, arch_of

ROIS_PER_IMAGE = 300   # cfg.TEST.RPN_POST_NMS_TOP_N (lib/mnc_config.py:130)
PRE_NMS_TOP_N = 6000   # cfg.TEST.RPN_PRE_NMS_TOP_N (:128)
RPN_NMS_THRESH = 0.7   # cfg.TEST.RPN_NMS_THRESH (:126)
RPN_MIN_SIZE = 16.0    # cfg.TEST.RPN_MIN_SIZE (:132)
MASK_SIZE = 21         # cfg.MASK_SIZE (:28)
NUM_CLASSES = 21


def _ceil_half(x):
    return (x + 1) // 2


class MNCEngine:
    def __init__(self, weights, device="cuda", impl="tc", sm_count=None):
        """weights: {caffe name: (weight, bias)} fp32 tensors in Caffe layouts (see weights.py)."""
        self.device = torch.device(device)
        self.impl = impl
        self.fuse_pool = True
        self.arch = arch_of(weights)
        self.sms = sm_count or torch.cuda.get_device_properties(self.device).multi_processor_count
        dev = self.device
        self.c5 = self.arch["trunk"][-1]
        self.fc = self.arch["fc"]
        self.me = self.arch["maskest"]
        w = {k: (v[0].to(dev).float().contiguous(), v[1].to(dev).float().contiguous())
             for k, v in weights.items()}
        # conv1_1 stays fp32 (SIMT kernel); the other convs are split [2, Cout, 9*Cin]
        self.conv1_1 = w["conv1_1"]
        # 64-channel conv1_1 runs on the tensor cores (stacked hi/lo weight tile, K padded to 32)
        self.conv1_1_tc = (dense.conv1_1_weight_to_tc(w["conv1_1"][0])
                           if impl == "tc" and w["conv1_1"][0].shape[0] == 64 else None)
        self.convs = []
        for name in TRUNK_NAMES[1:] + ["rpn_conv_3x3"]:
            if name in w:   # the CFM test net has no RPN (proposals are an input)
                self.convs.append((name, dense.conv_weight_to_split(w[name][0]), w[name][1]))
        self.trunk_convs = [c for c in self.convs if c[0] != "rpn_conv_3x3"]
        if "rpn_conv_3x3" in w:
            r = self.arch["rpn"]
            rpn_w = torch.cat([w["rpn_cls_score"][0].reshape(18, r), w["rpn_bbox_pred"][0].reshape(36, r)])
            self.rpn_head = (dense.split(rpn_w), torch.cat([w["rpn_cls_score"][1], w["rpn_bbox_pred"][1]]).contiguous())
        c5 = self.c5
        self.fc6 = (dense.fc_weight_to_split(w["fc6"][0], (c5, 7, 7)), w["fc6"][1])
        self.fc7 = (dense.split(w["fc7"][0]), w["fc7"][1])
        if "fc6_maskest" in w:
            self.fc6_maskest = (dense.fc_weight_to_split(w["fc6_maskest"][0], (c5, 14, 14)), w["fc6_maskest"][1])
            self.mask_pred = (dense.split(w["mask_pred"][0]), w["mask_pred"][1])
            self.fc6_mask = (dense.fc_weight_to_split(w["fc6_mask"][0], (c5, 7, 7)), w["fc6_mask"][1])
            self.fc7_mask = (dense.split(w["fc7_mask"][0]), w["fc7_mask"][1])
        # cls_score | seg_cls_score | bbox_pred share their input: one inner product for all of them
        names = [n for n in ("cls_score", "seg_cls_score", "bbox_pred") if n in w]
        self.cls_head_names = names
        cls_w = torch.cat([w[n][0] for n in names])
        cls_b = torch.cat([w[n][1] for n in names])
        self.cls_heads = (dense.split(cls_w), cls_b.contiguous())
        self._buf = {}

    # ------------------------------------------------------------------ helpers
    def _split_buf(self, key, *shape):
        t = self._buf.get(key)
        need = 1
        for s in shape:
            need *= s
        if t is None or t.numel() < 2 * need:
            t = torch.empty(2 * need, dtype=torch.bfloat16, device=self.device)
            self._buf[key] = t
        return t[:2 * need].view(2, *shape)

    def _f32_buf(self, key, *shape):
        t = self._buf.get(key)
        need = 1
        for s in shape:
            need *= s
        if t is None or t.numel() < need:
            t = torch.empty(need, dtype=torch.float32, device=self.device)
            self._buf[key] = t
        return t[:need].view(*shape)

    def _linear(self, a, M, K, wgt, N, bias, relu, out=None, out_f32=None, out_stride=None,
                out_ch_offset=0, key="lin", block_k=0, bn=0):
        """y = act(a @ W^T + b) through the implicit-GEMM kernel; split-K when the tile count
        cannot fill the GPU (e.g. fc6_maskest: K = 100352, N = 256)."""
        # Cout tile: 192 (BLOCK_K 32, 5 stages) for the wide layers -- at M = 2400, N = 4096 it gives
        # 19 x 22 = 418 tiles = 2.8 waves of 148 CTAs, against 4.1 (-> 5) waves at 128 and 2.05
        # (-> 3, at twice the tile cost) at 256; measured 480 vs 404 vs 347 TF/s on the fc6 shape
        # (profiles/r01_igemm_bk32_bn192.log)
        bn = bn or (64 if N <= 64 else (128 if N <= 128 else (192 if N >= 1024 else 256)))
        tiles = math.ceil(M / 128) * math.ceil(N / bn)
        k_steps = K // 64
        split = self._pick_split(tiles, k_steps) if self.impl == "tc" else 1
        a4 = a.view(2, 1, 1, M, K)
        if split == 1:
            dense.igemm(a4, 1, 1, M, K, wgt, N, 1, bias=bias, relu=relu, out=out, out_f32=out_f32,
                        out_pix_stride=out_stride, out_ch_offset=out_ch_offset, bn=bn,
                        impl=self.impl)
            return
        part = self._f32_buf("splitk_" + key, split, M, N)
        if block_k:
            dense.set_block_k(block_k)
        dense.igemm(a4, 1, 1, M, K, wgt, N, 1, out_f32=part, split_k=split, split_stride=M * N, bn=bn)
        if block_k:
            dense.set_block_k(0)
        dense.splitk_reduce(part, split, M * N, M, N, bias=bias, relu=relu, out=out,
                            out_f32=out_f32, out_row_stride=out_stride, out_ch_offset=out_ch_offset)

    def _pick_split(self, tiles, k_steps, max_split=32):
        """Split-K only when the launch cannot fill the GPU (e.g. fc6_maskest: 19 row tiles,
        K = 100352).  A wave-quantisation-driven split (608 tiles -> 4.1 waves) was measured and
        rejected: the extra fp32 partial reduce costs what the shorter tail saves
        (profiles/README.md).  When splitting, the factor minimises waves x k-steps-per-CTA:
        19 tiles x 8 = 152 work items on 148 SMs ran as two waves (0.47 ms); x 7 = 133 is one."""
        if tiles >= self.sms * 0.7 or k_steps < 16:
            return 1
        best, best_cost = 1, None
        for s in range(1, min(max_split, max(1, k_steps // 8)) + 1):
            waves = math.ceil(tiles * s / self.sms)
            cost = waves * math.ceil(k_steps / s) + 0.5 * s      # + partial-sum traffic per split
            if best_cost is None or cost < best_cost:
                best, best_cost = s, cost
        return best

    def _conv(self, x, B, H, W, cin, wgt, cout, bias, out, key):
        """3x3 conv + bias + ReLU -> split NHWC, split-K when whole waves would idle."""
        bn = 64 if cout <= 64 else (128 if cout <= 128 else 256)
        tiles = B * math.ceil(H / 8) * math.ceil(W / 16) * math.ceil(cout / bn)
        split = self._pick_split(tiles, 9 * cin // 64, max_split=4) if self.impl == "tc" else 1
        if split == 1:
            dense.igemm(x, B, H, W, cin, wgt, cout, 9, bias=bias, relu=True, out=out, impl=self.impl)
            return
        M = B * H * W
        part = self._f32_buf("splitk_" + key, split, M, cout)
        dense.igemm(x, B, H, W, cin, wgt, cout, 9, out_f32=part, split_k=split, split_stride=M * cout)
        dense.splitk_reduce(part, split, M * cout, M, cout, bias=bias, relu=True, out=out)

    # ------------------------------------------------------------------ trunk
    def trunk(self, data):
        """conv1_1 .. conv5_3 (test.prototxt:19-387).  data fp32 (B,3,H,W) -> split NHWC conv5_3."""
        B, _, H, W = data.shape
        ch = self.arch["trunk"]
        big = B * H * W * max(ch[0], ch[1])
        bufs = [self._split_buf("act0", big), self._split_buf("act1", big)]
        cur = 0
        x = bufs[cur].view(-1)[:2 * B * H * W * ch[0]].view(2, B, H, W, ch[0])
        if self.conv1_1_tc is not None:
            dense.conv1_1_tc(data.contiguous(), self.conv1_1_tc, self.conv1_1[1], x)
        else:
            dense.conv1_1(data, self.conv1_1[0], self.conv1_1[1], x)
        cin = ch[0]
        if "conv1_1" in POOL_AFTER:
            raise NotImplementedError
        for (name, wgt, bias) in self.trunk_convs:
            cout = wgt.shape[1]
            nxt = 1 - cur
            if name == "conv5_3":
                y = self._split_buf("conv5_3", B, H, W, cout)
            else:
                y = bufs[nxt].view(-1)[:2 * B * H * W * cout].view(2, B, H, W, cout)
            fuse = (name in POOL_AFTER) and self.impl == "tc" and self.fuse_pool
            if fuse:
                # Pooling fused into the conv epilogue: the full-resolution map is never written
                Ho, Wo = _ceil_half(H), _ceil_half(W)
                y = bufs[nxt].view(-1)[:2 * B * Ho * Wo * cout].view(2, B, Ho, Wo, cout)
                dense.igemm(x, B, H, W, cin, wgt, cout, 9, bias=bias, relu=True, out=y, pool=True)
                x, cur, cin, H, W = y, nxt, cout, Ho, Wo
                continue
            self._conv(x, B, H, W, cin, wgt, cout, bias, y, "conv")
            x, cur, cin = y, nxt, cout
            if name in POOL_AFTER:
                Ho, Wo = _ceil_half(H), _ceil_half(W)
                nxt = 1 - cur
                y = bufs[nxt].view(-1)[:2 * B * Ho * Wo * cout].view(2, B, Ho, Wo, cout)
                dense.maxpool2x2(x, B, H, W, cout, y)
                x, cur, H, W = y, nxt, Ho, Wo
        return x, H, W

    # ------------------------------------------------------------------ one cascade stage head
    def head(self, feat14, box7, R, tag):
        """test.prototxt:509-785 on R RoIs.  feat14 split [2,R,14,14,C5], box7 split [2,R,7,7,C5]."""
        c5, fc, me = self.c5, self.fc, self.me
        h_me = self._split_buf("h_me", R, me)
        # fc6_maskest streams its 963 MB activation matrix from HBM exactly once (a single Cout
        # tile: no L2 reuse), so it wants loads in flight rather than big stages: BLOCK_K 32 gives a
        # 4-deep ring at BN 256 (2-deep at 64 left every k-step waiting ~2 us for DRAM).  A 6-deep
        # ring (BN 128) is no faster: what remains (2.7 TB/s) is the DRAM efficiency of 128-byte
        # row segments 200 KB apart, the price of K-major rows with K = 100352.
        self._linear(feat14, R, 196 * c5, self.fc6_maskest[0], me, self.fc6_maskest[1], True,
                     out=h_me, key="me", block_k=32)
        logits = self._f32_buf("mask_logits_" + tag, R, 448)
        self._linear(h_me, R, me, self.mask_pred[0], 441, self.mask_pred[1], False,
                     out_f32=logits, out_stride=448, key="mp")
        mask_proposal, mask14 = ops.sigmoid_mask_resize(logits, R, MASK_SIZE, 14)
        join = self._split_buf("join", R, 2 * fc)
        h6 = self._split_buf("h6", R, fc)
        self._linear(box7, R, 49 * c5, self.fc6[0], fc, self.fc6[1], True, out=h6, key="fc6")
        self._linear(h6, R, fc, self.fc7[0], fc, self.fc7[1], True, out=join, out_stride=2 * fc,
                     out_ch_offset=fc, key="fc7")
        m7 = self._split_buf("m7", R, 7, 7, c5)
        ops.mask_pool_split(feat14, mask14, R, c5, m7)
        self._linear(m7, R, 49 * c5, self.fc6_mask[0], fc, self.fc6_mask[1], True, out=h6, key="fc6")
        self._linear(h6, R, fc, self.fc7_mask[0], fc, self.fc7_mask[1], True, out=join,
                     out_stride=2 * fc, out_ch_offset=0, key="fc7")
        heads = torch.empty((R, 128), dtype=torch.float32, device=self.device)
        self._linear(join, R, 2 * fc, self.cls_heads[0], 126, self.cls_heads[1], False,
                     out_f32=heads, out_stride=128, key="cls")
        cls_prob = ops.softmax_rows(heads[:, 0:21], 21)
        seg_cls_prob = ops.softmax_rows(heads[:, 21:42], 21)
        bbox_pred = heads[:, 42:126]
        return dict(mask_proposal=mask_proposal, mask_logits=logits, mask_resize=mask14,
                    cls_prob=cls_prob, seg_cls_prob=seg_cls_prob, bbox_pred=bbox_pred,
                    seg_cls_score=heads[:, 21:42], join=join)

    # ------------------------------------------------------------------ trunk + RPN + proposals
    def conv5_f32(self, conv5_3, B, H5, W5):
        """fp32 copy of conv5_3 (= hi + lo, exact) for the RoI gathers: 39 MB per batch of 8."""
        c5f = self._f32_buf("conv5_f32", B, H5, W5, self.c5)
        dense.split_to_f32(conv5_3, c5f)
        return c5f

    def rpn_rois(self, data, im_info, keep_intermediate=False):
        """test.prototxt:19-476: trunk, rpn_conv_3x3, rpn_cls_score | rpn_bbox_pred, softmax,
        ProposalLayer.  -> conv5_3 (split NHWC), H5, W5, fp32 conv5_3, rois (B*300,5), counts."""
        B = data.shape[0]
        conv5_3, H5, W5 = self.trunk(data)
        c5, r = self.c5, self.arch["rpn"]
        name, wgt, bias = self.convs[-1]
        rpn = self._split_buf("rpn", B, H5, W5, r)
        self._conv(conv5_3, B, H5, W5, c5, wgt, r, bias, rpn, "conv")
        rpn_out = self._f32_buf("rpn_out", B, H5, W5, 64)
        self._linear(rpn, B * H5 * W5, r, self.rpn_head[0], 54, self.rpn_head[1], False,
                     out_f32=rpn_out, out_stride=64, key="rpn")
        res = ops.proposals_from_rpn(rpn_out, None, im_info, B, H5, W5, "nhwc", True,
                                     pre_nms_top_n=PRE_NMS_TOP_N, post_nms_top_n=ROIS_PER_IMAGE,
                                     nms_thresh=RPN_NMS_THRESH, min_size=RPN_MIN_SIZE,
                                     batch_index_mode=True, return_intermediate=keep_intermediate)
        rois = res[0].view(B * ROIS_PER_IMAGE, 5)
        return conv5_3, H5, W5, self.conv5_f32(conv5_3, B, H5, W5), rois, res[1], res, rpn_out

    # ------------------------------------------------------------------ whole forward
    def forward(self, data, im_info, keep_intermediate=False):
        """data fp32 (B,3,H,W) device, im_info fp32 (B,3) device [h, w, scale].
        Returns device tensors named after the blobs callers read (tools/demo.py:84-90):
        rois (B*300,5), mask_proposal (B*300,1,21,21), seg_cls_prob (B*300,21) and the `_ext`
        versions, plus roi_counts (B,) = number of real (non-padding) RoIs per image."""
        B = data.shape[0]
        out = {}
        conv5_3, H5, W5, c5f, rois, roi_counts, res, rpn_out = self.rpn_rois(data, im_info, keep_intermediate)
        c5 = self.c5
        R = B * ROIS_PER_IMAGE
        out["rois"] = rois
        out["roi_counts"] = roi_counts
        feat14 = self._split_buf("feat14", R, 14, 14, c5)
        box7 = self._split_buf("box7", R, 7, 7, c5)
        ops.roi_warp_split(c5f, c5, H5, W5, rois, 2, feat14, box7)
        s1 = self.head(feat14, box7, R, "s1")
        rois_ext = ops.stage_bridge(rois, s1["bbox_pred"], s1["seg_cls_prob"], im_info,
                                    ROIS_PER_IMAGE)
        out["rois_ext"] = rois_ext
        for k in ("mask_proposal", "seg_cls_prob", "cls_prob", "bbox_pred"):
            out[k] = s1[k]
        if keep_intermediate:
            out["_rpn_out"] = rpn_out.clone()
            out["_proposal"] = res[2]
            out["_conv5_3"] = conv5_3.clone()
            out["_feat14"] = feat14.clone()
            out["_box7"] = box7.clone()
            out["_mask_logits"] = s1["mask_logits"].clone()
            out["_mask_resize"] = s1["mask_resize"]
            out["_join"] = s1["join"].clone()
        ops.roi_warp_split(c5f, c5, H5, W5, rois_ext, 1, feat14, box7)
        s2 = self.head(feat14, box7, R, "s2")
        for k in ("mask_proposal", "seg_cls_prob", "cls_prob", "bbox_pred"):
            out[k + "_ext"] = s2[k]
        if keep_intermediate:
            out["_feat14_ext"] = feat14.clone()
            out["_mask_logits_ext"] = s2["mask_logits"].clone()
        return out

    def detect(self, data, im_info, im_hw, im_scale):
        """forward + im_detect tail (tools/demo.py:92-100): boxes (B,600,4), masks (B,600,1,21,21),
        scores (B,600,21), valid (B,600) uint8."""
        o = self.forward(data, im_info)
        return self.detect_tail(o, data.shape[0], im_hw, im_scale) + (o,)

    def detect_tail(self, o, B, im_hw, im_scale, n=ROIS_PER_IMAGE):
        """The `im_detect` tail on the blobs of `forward` (tools/demo.py:84-100 ==
        TesterWrapper.py:244-260): rois / im_scale (fp32 division, the numpy-1.x evaluation of
        `rois[:, 1:5] / im_scales[0]`), clip_boxes to the ORIGINAL image shape im_hw, stage 1 rows
        then stage 2 rows.  o: dict with rois, rois_ext (B*n,5), mask_proposal(_ext),
        seg_cls_prob(_ext), roi_counts (B,); n rows per image and stage."""
        b1 = ops.unscale_clip(o["rois"], n, im_scale, im_hw).view(B, n, 4)
        b2 = ops.unscale_clip(o["rois_ext"], n, im_scale, im_hw).view(B, n, 4)
        boxes = torch.cat([b1, b2], dim=1).contiguous()
        masks = torch.cat([o["mask_proposal"].view(B, n, 1, MASK_SIZE, MASK_SIZE),
                           o["mask_proposal_ext"].view(B, n, 1, MASK_SIZE, MASK_SIZE)], dim=1).contiguous()
        scores = torch.cat([o["seg_cls_prob"].view(B, n, NUM_CLASSES),
                            o["seg_cls_prob_ext"].view(B, n, NUM_CLASSES)], dim=1).contiguous()
        ar = torch.arange(n, device=self.device, dtype=torch.int32).view(1, n)
        v = (ar < o["roi_counts"].view(B, 1)).to(torch.uint8)
        valid = torch.cat([v, v], dim=1).contiguous()
        return boxes, masks, scores, valid
