"""caffe.Net for the MNC 5-stage test graph, on the fused B200 engine."""
from collections import OrderedDict

import numpy as np


def load_weights(prototxt=None, weights=None):
    """Resolve the `weights` argument of caffe.Net into the engine's {layer: (weight, bias)} dict
    (and check the prototxt against the built-in 5-stage graph)."""
    import torch
    from mnc_b200.weights import make_weights
    from . import mnc_graph
    if prototxt is not None:
        mnc_graph.check_prototxt(prototxt)
    if weights is None:
        return make_weights()
    if isinstance(weights, str) and weights.endswith(".caffemodel"):
        from mnc_b200.caffemodel import weights_from_caffemodel
        return weights_from_caffemodel(weights)      # binary NetParameter, by layer name
    if isinstance(weights, str):
        return torch.load(weights, map_location="cpu")
    return weights


class Net(object):
    """Net(prototxt, weights, phase).

    prototxt : path to models/VGG16/mnc_5stage/test.prototxt (checked against the built-in graph)
               or None for the built-in graph.
    weights  : {caffe layer name: (weight, bias)} dict, a binary `.caffemodel`
               (mnc_b200/caffemodel.py), a torch-saved file of such a dict, or None for the seeded
               random initialiser (mnc_b200/weights.py).  `.caffemodel.h5` needs h5py (absent).
    """

    def __init__(self, prototxt=None, weights=None, phase=1):
        import torch
        from mnc_b200.engine import MNCEngine
        from . import Blob, _state, mnc_graph
        if phase != 1:
            raise NotImplementedError("inference (caffe.TEST) only")
        weights = load_weights(prototxt, weights)
        self._graph = mnc_graph.build_graph()
        self._device = torch.device("cuda", _state["device"])
        with torch.cuda.device(self._device):
            self._engine = MNCEngine(weights, device=self._device)
        self.name = "VGG16"
        self.inputs = ["data", "im_info"]
        self.blobs = OrderedDict()
        self.blobs["data"] = Blob(1, 3, 224, 224)     # input_shape of test.prototxt:3-9
        self.blobs["im_info"] = Blob(1, 3)            # :11-15
        for layer in self._graph:
            for t in layer["top"]:
                if t not in self.blobs:
                    self.blobs[t] = Blob()
        self.outputs = ["cls_prob", "cls_prob_ext", "bbox_pred_ext"]

    def forward(self, blobs=None, start=None, end=None, **kwargs):
        """pycaffe.py:62-108: inputs by keyword, returns {output blob name: array}."""
        import torch
        from mnc_b200 import dense
        if start is not None or end is not None:
            raise NotImplementedError("partial forward is not supported by the fused engine")
        if kwargs:
            if set(kwargs.keys()) != set(self.inputs):
                raise Exception("Input blob arguments do not match net inputs.")
            for in_, blob in kwargs.items():
                if blob.shape[0] != self.blobs[in_].num:
                    raise Exception("Input is not batch sized")
                self.blobs[in_].data[...] = blob
        data = self.blobs["data"].data
        im_info = self.blobs["im_info"].data
        if data.shape[0] != 1:
            raise AssertionError("Only single item batches are supported")  # proposal_layer.py:65
        dev = self._device
        with torch.cuda.device(dev):
            d = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(dev)
            info = torch.from_numpy(np.ascontiguousarray(im_info, dtype=np.float32)).to(dev)
            o = self._engine.forward(d, info, keep_intermediate=True)
            n = int(o["roi_counts"][0].item())
            conv5 = o["_conv5_3"]
            _, B, H5, W5, C5 = conv5.shape
            c5 = torch.empty((B, C5, H5, W5), dtype=torch.float32, device=dev)
            dense.split_to_nchw(conv5, B, H5, W5, C5, c5)
            host = {
                "conv5_3": c5,
                "rois": o["rois"][:n], "rois_ext": o["rois_ext"][:n],
                "mask_proposal": o["mask_proposal"][:n], "mask_proposal_ext": o["mask_proposal_ext"][:n],
                "seg_cls_prob": o["seg_cls_prob"][:n], "seg_cls_prob_ext": o["seg_cls_prob_ext"][:n],
                "cls_prob": o["cls_prob"][:n], "cls_prob_ext": o["cls_prob_ext"][:n],
                "bbox_pred": o["bbox_pred"][:n], "bbox_pred_ext": o["bbox_pred_ext"][:n],
            }
            for k, v in host.items():
                arr = v.contiguous().cpu().numpy()
                if k.startswith("rois"):
                    arr[:, 0] = 0  # single-image net: batch index 0 (proposal_layer.py:159)
                self.blobs[k].data = arr
        return {k: self.blobs[k].data for k in self.outputs}
