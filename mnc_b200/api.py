"""Public host-buffer API of the batched engine: the call a user makes.

`Detector.im_detect_batch(host_blob)` is the batched form of the reference's `im_detect`
(tools/demo.py:79-100 == lib/caffeWrapper/TesterWrapper.py:239-260): network inputs in HOST memory
(fp32 NCHW blobs exactly as `prepare_mnc_args` builds them, tools/demo.py:54-76), results back in
HOST memory -- boxes (B,600,4), masks (B,600,1,21,21), scores (B,600,21), valid (B,600).
Host<->device copies go through pinned staging buffers on the engine's stream.
`Detector.mask_voting` is the batched `gpu_mask_voting` (lib/transform/mask_transform.py:213-286).
"""
import numpy as np
import torch

from . import ops
from .engine import MNCEngine, ROIS_PER_IMAGE, MASK_SIZE, NUM_CLASSES


class Detector:
    def __init__(self, weights, device="cuda", max_batch=8, height=600, width=1000, use_graph=True):
        self.device = torch.device(device)
        self.engine = MNCEngine(weights, device=self.device)
        # a step is ~60 launches on static buffers: replayed from a CUDA graph per input shape
        self.use_graph = use_graph
        self.max_batch = max_batch
        B, n = max_batch, 2 * ROIS_PER_IMAGE
        self._h_in = torch.empty((B, 3, height, width), dtype=torch.float32).pin_memory()
        # results come back as ONE record (ops.record_layout) + the valid flags: two D2H copies
        self._h_rec = torch.empty(ops.record_layout(B, ROIS_PER_IMAGE)[3], dtype=torch.float32).pin_memory()
        self._h_valid = torch.empty((B, n), dtype=torch.uint8).pin_memory()
        self._d_in = torch.empty((B, 3, height, width), dtype=torch.float32, device=self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _fit_input(self, H, W):
        """Input staging follows the blob size (real images scale to 600x800 ... 901x600,
        lib/utils/blob.py:41-46); the engine's own buffers grow on demand."""
        if tuple(self._d_in.shape[2:]) != (H, W):
            B = self.max_batch
            self._h_in = torch.empty((B, 3, H, W), dtype=torch.float32).pin_memory()
            self._d_in = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device)

    def im_detect_batch(self, blob, im_info=None, im_scales=None, im_shapes=None):
        """blob: (B,3,H,W) fp32 numpy / CPU tensor (mean-subtracted BGR, as `im_list_to_blob`
        returns).  im_info: (B,3) [H, W, scale] (default: blob size, scale 1).  Synchronous."""
        blob = torch.as_tensor(blob)
        B, _, H, W = blob.shape
        assert B <= self.max_batch
        self._fit_input(H, W)
        dev = self.device
        if im_info is None:
            im_info = np.tile(np.array([[H, W, 1.0]], dtype=np.float32), (B, 1))
        info_h = torch.as_tensor(np.asarray(im_info, dtype=np.float32))
        scale_h = info_h[:, 2].contiguous() if im_scales is None else torch.as_tensor(np.asarray(im_scales, np.float32))
        if im_shapes is None:
            # boxes are clipped to the ORIGINAL image (TesterWrapper.py:254-255): blob size / scale
            hw_h = torch.round(info_h[:, :2] / scale_h[:, None]).contiguous()
        else:
            hw_h = torch.as_tensor(np.asarray(im_shapes, np.float32))
        if blob.is_pinned():
            src = blob                                   # caller already handed page-locked memory
        else:
            self._h_in[:B].copy_(blob)                   # user memory -> pinned staging
            src = self._h_in[:B]
        with torch.cuda.device(dev):
            self._d_in[:B].copy_(src, non_blocking=True)
            info = info_h.to(dev, non_blocking=True)
            boxes, masks, scores, valid, _ = self._detect(
                self._d_in[:B], info, hw_h.to(dev), scale_h.to(dev))
            out = self._results_to_host(B, valid)
        self.h2d_bytes = blob.numel() * 4 + info_h.numel() * 4 + scale_h.numel() * 4 + hw_h.numel() * 4
        return out

    def _detect(self, data, info, hw, sc):
        # every 32nd call: did any activation outgrow the exponents frozen at calibration?
        self._calls = getattr(self, "_calls", 0) + 1
        if self._calls % 32 == 0:
            self.engine.range_ok()
        if self.use_graph:
            return self.engine.detect_graphed(data, info, hw, sc)
        return self.engine.detect(data, info, hw, sc)

    def _results_to_host(self, B, valid):
        rec = self.engine.last_record
        n = rec.numel()
        self._h_rec[:n].copy_(rec, non_blocking=True)
        self._h_valid[:B].copy_(valid, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.d2h_bytes = n * 4 + valid.numel()
        _, boxes, scores, masks = ops.record_views(self._h_rec[:n], B, ROIS_PER_IMAGE)
        return boxes.numpy(), masks.numpy(), scores.numpy(), self._h_valid[:B].numpy()

    def im_detect_images(self, images_u8):
        """Batched `im_detect(im, net)` on raw images, as the reference's callers hand them over
        (tools/demo.py:143-146): uint8 BGR (B,H,W,3) host array, all of one size.  Mean
        subtraction, the 600/1000 resize rule and HWC->NCHW run on the device (mnc_prep_images),
        so only B*H*W*3 bytes cross PCIe.  Returns boxes (in original-image coordinates), masks,
        scores, valid -- host arrays -- and the scale used."""
        pinned_src = None
        if isinstance(images_u8, torch.Tensor):
            # a page-locked uint8 tensor goes to the device without the staging copy
            assert images_u8.dtype == torch.uint8 and images_u8.is_contiguous()
            if images_u8.is_pinned():
                pinned_src = images_u8
            images_u8 = images_u8.numpy()
        images_u8 = np.ascontiguousarray(images_u8)
        B, H, W, _ = images_u8.shape
        scale = ops.im_scale_for((H, W))
        out_h, out_w = int(np.rint(H * scale)), int(np.rint(W * scale))
        assert B <= self.max_batch
        self._fit_input(out_h, out_w)
        dev = self.device
        if getattr(self, "_h_u8", None) is None or self._h_u8.shape[1:] != images_u8.shape[1:]:
            self._h_u8 = torch.empty((self.max_batch, H, W, 3), dtype=torch.uint8).pin_memory()
            self._d_u8 = torch.empty((self.max_batch, H, W, 3), dtype=torch.uint8, device=dev)
        if pinned_src is None:
            self._h_u8[:B].copy_(torch.from_numpy(images_u8))
            pinned_src = self._h_u8[:B]
        info = torch.tensor([[out_h, out_w, scale]] * B, dtype=torch.float32)
        hw = torch.tensor([[H, W]] * B, dtype=torch.float32)
        sc = torch.full((B,), scale, dtype=torch.float32)
        with torch.cuda.device(dev):
            self._d_u8[:B].copy_(pinned_src, non_blocking=True)
            ops.prep_images(self._d_u8[:B], scale, out=self._d_in[:B])
            boxes, masks, scores, valid, _ = self._detect(
                self._d_in[:B], info.to(dev, non_blocking=True), hw.to(dev, non_blocking=True),
                sc.to(dev, non_blocking=True))
            out = self._results_to_host(B, valid)
        self.h2d_bytes = images_u8.nbytes + (info.numel() + hw.numel() + sc.numel()) * 4
        return out + (scale,)

    def im_detect_stream(self, batches):
        """Pipelined `im_detect_images`: an iterable of uint8 BGR (B,H,W,3) host batches (all of
        one size) -> a generator of (boxes, masks, scores, valid, scale) per batch, in order.
        Two batches are in flight: while batch k computes, the frames of batch k+1 cross PCIe on a
        copy stream and the record of batch k-1 comes back on another, so the host<->device copies
        (14.4 MB in, 9 MB out per batch of 8) leave the critical path; and the two batches compute
        on two streams, each with its own engine state over the shared weights
        (MNCEngine.clone_state), so that batch k+1's kernels fill batch k's wave tails and its
        low-occupancy proposal phase.  A yielded result is valid until the next iteration (its
        pinned buffers are reused two batches later)."""
        dev = self.device
        if getattr(self, "_s_in", None) is None:
            self._s_in = torch.cuda.Stream(device=dev)
            self._s_out = torch.cuda.Stream(device=dev)
            self._slots = [None, None]
            self._s_comp = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
            self._engines = [self.engine, None]
        pending = None
        for k, images_u8 in enumerate(batches):
            cur = self._submit(k & 1, images_u8)
            if pending is not None:
                yield self._collect(pending)
            pending = cur
        if pending is not None:
            yield self._collect(pending)

    def _submit(self, slot, images_u8):
        dev = self.device
        pinned_src = None
        if isinstance(images_u8, torch.Tensor):
            assert images_u8.dtype == torch.uint8 and images_u8.is_contiguous()
            if images_u8.is_pinned():
                pinned_src = images_u8
            images_u8 = images_u8.numpy()
        images_u8 = np.ascontiguousarray(images_u8)
        B, H, W, _ = images_u8.shape
        assert B <= self.max_batch
        scale = ops.im_scale_for((H, W))
        out_h, out_w = int(np.rint(H * scale)), int(np.rint(W * scale))
        self._fit_input(out_h, out_w)
        st = self._slots[slot]
        if st is None or tuple(st["h_u8"].shape[1:]) != (H, W, 3):
            n_rec = ops.record_layout(self.max_batch, ROIS_PER_IMAGE)[3]
            st = dict(h_u8=torch.empty((self.max_batch, H, W, 3), dtype=torch.uint8).pin_memory(),
                      d_u8=torch.empty((self.max_batch, H, W, 3), dtype=torch.uint8, device=dev),
                      d_rec=torch.empty(n_rec, dtype=torch.float32, device=dev),
                      h_rec=torch.empty(n_rec, dtype=torch.float32).pin_memory(),
                      d_in=torch.empty_like(self._d_in),
                      u8_free=None, out_done=None)
            self._slots[slot] = st
        if tuple(st["d_in"].shape) != tuple(self._d_in.shape):
            st["d_in"] = torch.empty_like(self._d_in)
        eng = self._engines[slot]
        if eng is None:          # slot 1's engine: made once slot 0's first call has calibrated
            torch.cuda.synchronize(dev)
            eng = self._engines[slot] = self.engine.clone_state()
        if pinned_src is None:
            st["h_u8"][:B].copy_(torch.from_numpy(images_u8))      # pageable -> pinned staging (host)
            pinned_src = st["h_u8"][:B]
        info = torch.tensor([[out_h, out_w, scale]] * B, dtype=torch.float32)
        hw = torch.tensor([[H, W]] * B, dtype=torch.float32)
        sc = torch.full((B,), scale, dtype=torch.float32)
        with torch.cuda.device(dev), torch.cuda.stream(self._s_comp[slot]):
            main = torch.cuda.current_stream()                      # this slot's compute stream
            with torch.cuda.stream(self._s_in):                     # frames of this batch: H2D
                if st["u8_free"] is not None:
                    self._s_in.wait_event(st["u8_free"])
                st["d_u8"][:B].copy_(pinned_src, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self._s_in)
            main.wait_event(ev_in)
            ops.prep_images(st["d_u8"][:B], scale, out=st["d_in"][:B])
            st["u8_free"] = torch.cuda.Event()
            st["u8_free"].record(main)
            if st["out_done"] is not None:
                main.wait_event(st["out_done"])                     # this slot's record was read
            n = ops.record_layout(B, ROIS_PER_IMAGE)[3]
            args = (st["d_in"][:B], info.to(dev, non_blocking=True), hw.to(dev, non_blocking=True),
                    sc.to(dev, non_blocking=True))
            self._calls = getattr(self, "_calls", 0) + 1
            if self._calls % 32 in (0, 1) and self._calls > 1:
                if not eng.range_ok():                              # exponents are shared: both re-measure
                    for e in self._engines:
                        if e is not None:
                            e._calibrated = False
                            e._graphs.clear() if hasattr(e, "_graphs") else None
            if self.use_graph:
                eng.detect_graphed(*args, rec=st["d_rec"])
            else:
                o = eng.forward(args[0], args[1])
                eng.detect_tail(o, B, args[2], args[3], rec=st["d_rec"])
            ev_done = torch.cuda.Event()
            ev_done.record(main)
            with torch.cuda.stream(self._s_out):                    # record of this batch: D2H
                self._s_out.wait_event(ev_done)
                st["h_rec"][:n].copy_(st["d_rec"][:n], non_blocking=True)
                st["out_done"] = torch.cuda.Event()
                st["out_done"].record(self._s_out)
        self.h2d_bytes = images_u8.nbytes + (info.numel() + hw.numel() + sc.numel()) * 4
        self.d2h_bytes = n * 4
        return (slot, B, n, scale)

    def _collect(self, handle):
        slot, B, n, scale = handle
        st = self._slots[slot]
        st["out_done"].synchronize()
        counts, boxes, scores, masks = ops.record_views(st["h_rec"][:n], B, ROIS_PER_IMAGE)
        # valid flags from the counts (2 x RoIs per image: stage-1 rows, then stage-2 rows)
        per_stage = (counts.numpy() / 2).astype(np.int64)
        idx = np.arange(2 * ROIS_PER_IMAGE) % ROIS_PER_IMAGE
        valid = (idx[None, :] < per_stage[:, None]).astype(np.uint8)
        return boxes.numpy(), masks.numpy(), scores.numpy(), valid, scale

    def mask_voting(self, boxes, masks, scores, valid, im_hw, max_per_image=100):
        """Device-resident batched gpu_mask_voting on `engine.detect` outputs (device tensors)."""
        hw = torch.as_tensor(np.asarray(im_hw, dtype=np.int32)).to(self.device)
        return ops.mask_voting_checked(boxes, masks, scores, hw, max_per_image=max_per_image,
                                       box_valid=valid)
