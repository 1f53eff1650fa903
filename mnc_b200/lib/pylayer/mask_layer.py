"""pylayer.mask_layer.MaskLayer -- reference lib/pylayer/mask_layer.py:95-102 (forward_test is a
pure reshape (N,441) -> (N,1,21,21); the fused engine never leaves the device for it)."""
import caffe
from mnc_config import cfg


class MaskLayer(caffe.Layer):
    def setup(self, bottom, top):
        top[0].reshape(1, 1, cfg.MASK_SIZE, cfg.MASK_SIZE)

    def forward(self, bottom, top):
        mask_pred = bottom[0].data
        n = mask_pred.shape[0]
        out = mask_pred.reshape((n, 1, cfg.MASK_SIZE, cfg.MASK_SIZE))
        top[0].reshape(*out.shape)
        top[0].data[...] = out
