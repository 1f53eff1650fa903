"""pylayer.proposal_layer.ProposalLayer -- reference lib/pylayer/proposal_layer.py:21-175 (TEST).

Same caffe.Layer protocol (param_str_ YAML with feat_stride, setup/reshape/forward, top reshaped
inside forward), but the body runs on the device: decode + clip + min-size filter, rank sort,
top-6000, bitmask NMS with device-side scan, top-300."""
import numpy as np
import torch
import yaml

import caffe
from mnc_config import cfg
from mnc_b200 import ops


class ProposalLayer(caffe.Layer):
    def setup(self, bottom, top):
        layer_params = yaml.safe_load(self.param_str_) if self.param_str_ else {}
        self._feat_stride = layer_params.get("feat_stride", 16)
        self._num_anchors = 9
        top[0].reshape(1, 5)

    def reshape(self, bottom, top):
        """Reshaping happens during the call to forward."""
        pass

    def forward(self, bottom, top):
        assert bottom[0].data.shape[0] == 1, 'Only single item batches are supported'
        cfg_key = str(self.phase)
        if cfg_key != "TEST":
            raise NotImplementedError("training path is out of scope")
        c = cfg[cfg_key]
        dev = torch.device("cuda", cfg.GPU_ID)
        cls = torch.from_numpy(np.ascontiguousarray(bottom[0].data, dtype=np.float32)).to(dev)
        bbox = torch.from_numpy(np.ascontiguousarray(bottom[1].data, dtype=np.float32)).to(dev)
        im_info = torch.from_numpy(np.ascontiguousarray(bottom[2].data, dtype=np.float32)).to(dev)
        H, W = cls.shape[-2:]
        with torch.cuda.device(dev):
            rois, counts = ops.proposals_from_rpn(
                cls, bbox, im_info.view(-1, 3), 1, H, W, "nchw", apply_softmax=False,
                pre_nms_top_n=c.RPN_PRE_NMS_TOP_N, post_nms_top_n=c.RPN_POST_NMS_TOP_N,
                nms_thresh=c.RPN_NMS_THRESH, min_size=float(c.RPN_MIN_SIZE),
                batch_index_mode=False)
            n = int(counts[0].item())
            blob = rois[0, :n].cpu().numpy()
        top[0].reshape(*blob.shape)
        top[0].data[...] = blob
