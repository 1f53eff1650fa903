"""Reading and writing binary `.caffemodel` files ("next" row 2).

The reference loads trained weights by layer name from a serialized `NetParameter`
(caffe-mnc/src/caffe/net.cpp:837-845 -> CopyTrainedLayersFromBinaryProto; shared parameters are
matched by `param{name}`, net.cpp:470-509).  There is no protobuf schema compiler in this image,
so the few message fields needed are decoded straight from the protobuf wire format, field numbers
from caffe-mnc/src/caffe/proto/caffe.proto:
    NetParameter   : name = 1, layer = 100 (LayerParameter), layers = 2 (V1LayerParameter)
    LayerParameter : name = 1, type = 2, blobs = 7 (BlobProto)      (:311-329)
    V1LayerParameter: name = 4, blobs = 6
    BlobProto      : shape = 7 (BlobShape), data = 5 (packed float), double_data = 8,
                     legacy dims num/channels/height/width = 1..4    (:10-22)
    BlobShape      : dim = 1 (packed int64)                          (:6-8)
`.caffemodel.h5` (HDF5, what data/scripts/fetch_mnc_model.sh downloads; written by Net::ToHDF5,
net.cpp:920-975) is read by the small HDF5 parser in mnc_b200/hdf5_min.py -- no h5py needed.
"""
import struct

import numpy as np


# ----------------------------------------------------------------------------- wire format
def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) over one message; length-delimited values are
    memoryview slices."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fno, wt, v


def _packed_varints(buf):
    out, pos = [], 0
    while pos < len(buf):
        v, pos = _varint(buf, pos)
        out.append(v)
    return out


def _parse_blob(buf):
    dims, legacy, data = None, {}, None
    floats = []
    for fno, wt, v in _fields(buf):
        if fno == 7 and wt == 2:
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    dims = _packed_varints(v2) if w2 == 2 else (dims or []) + [v2]
        elif fno == 5:
            if wt == 2:
                data = np.frombuffer(bytes(v), dtype="<f4")
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 8 and wt == 2:
            data = np.frombuffer(bytes(v), dtype="<f8").astype(np.float32)
        elif fno in (1, 2, 3, 4) and wt == 0:
            legacy[fno] = v
    if data is None:
        data = np.asarray(floats, dtype=np.float32)
    if dims is None:
        dims = [legacy.get(i, 1) for i in (1, 2, 3, 4)] if legacy else [data.size]
    return np.array(data, dtype=np.float32).reshape(dims)


def load_caffemodel(path):
    """-> {layer name: [blob ndarray, ...]} for every layer that carries blobs."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    layers = {}
    for fno, wt, v in _fields(buf):
        if wt != 2 or fno not in (100, 2):
            continue
        name_field, blob_field = (1, 7) if fno == 100 else (4, 6)
        name, blobs = None, []
        for f2, w2, v2 in _fields(v):
            if f2 == name_field and w2 == 2:
                name = bytes(v2).decode("utf-8")
            elif f2 == blob_field and w2 == 2:
                blobs.append(_parse_blob(v2))
        if name is not None and blobs:
            layers[name] = blobs
    return layers


MNC_LAYERS = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1",
              "conv4_2", "conv4_3", "conv5_1", "conv5_2", "conv5_3", "rpn_conv_3x3", "rpn_cls_score",
              "rpn_bbox_pred", "fc6_maskest", "mask_pred", "fc6", "fc7", "fc6_mask", "fc7_mask",
              "cls_score", "seg_cls_score", "bbox_pred"]


_RPN = ("rpn_conv_3x3", "rpn_cls_score", "rpn_bbox_pred")
_MASK = ("fc6_maskest", "mask_pred", "fc6_mask", "fc7_mask", "seg_cls_score")
GRAPH_LAYERS = {
    "mnc_5stage": MNC_LAYERS,
    "faster_rcnn": [n for n in MNC_LAYERS if n not in _MASK],     # faster_rcnn_end2end/test.prototxt
    "cfm": [n for n in MNC_LAYERS if n not in _RPN],              # cfm/test.prototxt (no RPN)
}
# layer names that differ between the graphs' prototxts (faster_rcnn_end2end/test.prototxt:391-405)
_ALIASES = {"rpn_conv_3x3": ("rpn_conv/3x3",)}


def weights_from_caffemodel(path, kind="mnc_5stage"):
    """The engine's weight dict {name: (weight, bias)} from a caffemodel of one of the supported
    test graphs.  The `_ext` layers share the owners' parameters (test.prototxt:829-834 ...), so
    only the owners are read.  (The reference's snapshots already hold un-normalised bbox_pred
    weights, lib/caffeWrapper/SolverWrapper.py:67-115.)"""
    import torch
    if path.endswith((".h5", ".hdf5")):
        from .hdf5_min import load_caffemodel_h5
        layers = load_caffemodel_h5(path)      # {layer: [blobs]} from /data/<layer>/<index>
    else:
        layers = load_caffemodel(path)
    for name, others in _ALIASES.items():
        for o in others:
            if name not in layers and o in layers:
                layers[name] = layers[o]
    wanted = GRAPH_LAYERS[kind]
    missing = [n for n in wanted if n not in layers]
    if missing:
        raise KeyError("caffemodel lacks %s layers: %s" % (kind, ", ".join(missing)))
    out = {}
    for n in wanted:
        blobs = layers[n]
        w = blobs[0]
        b = blobs[1].reshape(-1) if len(blobs) > 1 else np.zeros(w.shape[0], np.float32)
        if w.ndim == 4 and not n.startswith(("conv", "rpn_")):
            w = w.reshape(w.shape[-2], w.shape[-1])   # legacy (1,1,N,K) InnerProduct blobs
        out[n] = (torch.from_numpy(np.array(w, dtype=np.float32, order="C")),
                  torch.from_numpy(np.array(b, dtype=np.float32, order="C")))   # owned, writable copies
    return out


# ----------------------------------------------------------------------------- writer
def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_ld(fno, payload):
    return _enc_varint((fno << 3) | 2) + _enc_varint(len(payload)) + payload


def _enc_blob(arr):
    arr = np.ascontiguousarray(arr, dtype="<f4")
    shape = _enc_ld(1, b"".join(_enc_varint(int(d)) for d in arr.shape))
    return _enc_ld(7, shape) + _enc_ld(5, arr.tobytes())


def save_caffemodel(weights, path, net_name="VGG16", layer_types=None):
    """Serialise {name: (weight, bias)} as a NetParameter with one LayerParameter per entry."""
    layer_types = layer_types or {}
    out = bytearray(_enc_ld(1, net_name.encode()))
    for name, (w, b) in weights.items():
        w = w.numpy() if hasattr(w, "numpy") else np.asarray(w)
        b = b.numpy() if hasattr(b, "numpy") else np.asarray(b)
        ltype = layer_types.get(name, "Convolution" if w.ndim == 4 else "InnerProduct")
        body = _enc_ld(1, name.encode()) + _enc_ld(2, ltype.encode())
        body += _enc_ld(7, _enc_blob(w)) + _enc_ld(7, _enc_blob(b))
        out += _enc_ld(100, body)
    with open(path, "wb") as f:
        f.write(bytes(out))
