"""pylayer.stage_bridge_layer.StageBridgeLayer -- reference
lib/pylayer/stage_bridge_layer.py:237-255 (forward_test), body on the device."""
import numpy as np
import torch

import caffe
from mnc_config import cfg
from mnc_b200 import ops


class StageBridgeLayer(caffe.Layer):
    def setup(self, bottom, top):
        top[0].reshape(1, 5)

    def forward(self, bottom, top):
        dev = torch.device("cuda", cfg.GPU_ID)
        t = lambda b: torch.from_numpy(np.ascontiguousarray(b.data, dtype=np.float32)).to(dev)
        rois, deltas, scores, im_info = t(bottom[0]), t(bottom[1]), t(bottom[2]), t(bottom[3])
        with torch.cuda.device(dev):
            out = ops.stage_bridge(rois.view(-1, 5), deltas, scores, im_info.view(-1, 3),
                                   rois.shape[0])
            blob = out.cpu().numpy()
        blob[:, 0] = 0
        top[0].reshape(*blob.shape)
        top[0].data[...] = blob
