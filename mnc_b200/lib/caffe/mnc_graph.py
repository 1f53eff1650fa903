"""The MNC 5-stage test graph as data, written from models/VGG16/mnc_5stage/test.prototxt
(layer names, types, bottoms, tops and the parameters the path consumes), plus a small prototxt
reader (no protoc in this image) used to check a user-supplied prototxt against it."""
import re


def _conv(name, bottom, top, n, k=3, pad=1):
    return dict(name=name, type="Convolution", bottom=[bottom], top=[top],
                convolution_param=dict(num_output=n, kernel_size=k, pad=pad))


def _relu(name, blob):
    return dict(name=name, type="ReLU", bottom=[blob], top=[blob])


def _pool(name, bottom, top):
    return dict(name=name, type="Pooling", bottom=[bottom], top=[top],
                pooling_param=dict(pool="MAX", kernel_size=2, stride=2))


def _ip(name, bottom, top, n):
    return dict(name=name, type="InnerProduct", bottom=[bottom], top=[top],
                inner_product_param=dict(num_output=n))


def build_graph():
    L = []
    cfgs = [("1", 2, 64), ("2", 2, 128), ("3", 3, 256), ("4", 3, 512), ("5", 3, 512)]
    prev = "data"
    for blk, n, ch in cfgs:
        for i in range(1, n + 1):
            nm = "conv%s_%d" % (blk, i)
            L.append(_conv(nm, prev, nm, ch))
            L.append(_relu("relu%s_%d" % (blk, i), nm))
            prev = nm
        if blk != "5":
            L.append(_pool("pool" + blk, prev, "pool" + blk))
            prev = "pool" + blk
    L.append(_conv("rpn_conv_3x3", "conv5_3", "rpn_output", 512))
    L.append(_relu("rpn_relu_3x3", "rpn_output"))
    L.append(_conv("rpn_cls_score", "rpn_output", "rpn_cls_score", 18, k=1, pad=0))
    L.append(_conv("rpn_bbox_pred", "rpn_output", "rpn_bbox_pred", 36, k=1, pad=0))
    L.append(dict(name="rpn_cls_score_reshape", type="Reshape", bottom=["rpn_cls_score"],
                  top=["rpn_cls_score_reshape"]))
    L.append(dict(name="rpn_cls_prob", type="Softmax", bottom=["rpn_cls_score_reshape"],
                  top=["rpn_cls_prob"]))
    L.append(dict(name="rpn_cls_prob_reshape", type="Reshape", bottom=["rpn_cls_prob"],
                  top=["rpn_cls_prob_reshape"]))
    L.append(dict(name="proposal", type="Python",
                  bottom=["rpn_cls_prob_reshape", "rpn_bbox_pred", "im_info"], top=["rois"],
                  python_param=dict(module="pylayer.proposal_layer", layer="ProposalLayer")))
    for ext in ("", "_ext"):
        if ext == "":
            L.append(dict(name="roi_interpolate_conv5_premax", type="ROIWarping",
                          bottom=["conv5_3", "rois"], top=["roi_interpolate_conv5_premax"],
                          roi_warping_param=dict(pooled_w=28, pooled_h=28, spatial_scale=0.0625)))
            L.append(_pool("roi_interpolate_conv5", "roi_interpolate_conv5_premax",
                           "roi_interpolate_conv5"))
        else:
            L.append(dict(name="roi_interpolate_conv5_ext", type="ROIWarping",
                          bottom=["conv5_3", "rois_ext"], top=["roi_interpolate_conv5_ext"],
                          roi_warping_param=dict(pooled_w=14, pooled_h=14, spatial_scale=0.0625)))
        feat = "roi_interpolate_conv5" + ext
        L.append(_ip("fc6_maskest" + ext, feat, "fc6_maskest" + ext, 256))
        L.append(_relu("relu6_maskest" + ext, "fc6_maskest" + ext))
        L.append(_ip("mask_pred" + ext, "fc6_maskest" + ext, "mask_pred" + ext, 441))
        L.append(dict(name="mask_output" + ext, type="Sigmoid", bottom=["mask_pred" + ext],
                      top=["mask_output" + ext]))
        L.append(dict(name="mask_proposal" + ext, type="Python", bottom=["mask_output" + ext],
                      top=["mask_proposal" + ext],
                      python_param=dict(module="pylayer.mask_layer", layer="MaskLayer")))
        L.append(dict(name="mask_resize" + ext, type="MaskResize", bottom=["mask_proposal" + ext],
                      top=["mask_proposal_resize" + ext],
                      mask_resize_param=dict(output_height=14, output_width=14)))
        L.append(_pool("roi_interpolate_conv5_box" + ext, feat, "roi_interpolate_conv5_box" + ext))
        L.append(_ip("fc6" + ext, "roi_interpolate_conv5_box" + ext, "fc6" + ext, 4096))
        L.append(_relu("relu6" + ext, "fc6" + ext))
        L.append(_ip("fc7" + ext, "fc6" + ext, "fc7" + ext, 4096))
        L.append(_relu("relu7" + ext, "fc7" + ext))
        L.append(dict(name="mask_pooling" + ext, type="MaskPooling",
                      bottom=[feat, "mask_proposal_resize" + ext], top=["roi_mask_conv5" + ext]))
        L.append(_pool("roi_interpolate_conv5_mask" + ext, "roi_mask_conv5" + ext,
                       "roi_interpolate_conv5_mask" + ext))
        L.append(_ip("fc6_mask" + ext, "roi_interpolate_conv5_mask" + ext, "fc6_mask" + ext, 4096))
        L.append(_relu("relu6_mask" + ext, "fc6_mask" + ext))
        L.append(_ip("fc7_mask" + ext, "fc6_mask" + ext, "fc7_mask" + ext, 4096))
        L.append(_relu("relu7_mask" + ext, "fc7_mask" + ext))
        L.append(dict(name="join_box_mask" + ext, type="Concat",
                      bottom=["fc7_mask" + ext, "fc7" + ext], top=["join_box_mask" + ext]))
        L.append(_ip("cls_score" + ext, "join_box_mask" + ext, "cls_score" + ext, 21))
        L.append(dict(name="cls_prob" + ext, type="Softmax", bottom=["cls_score" + ext],
                      top=["cls_prob" + ext]))
        L.append(_ip("seg_cls_score" + ext, "join_box_mask" + ext, "seg_cls_score" + ext, 21))
        L.append(dict(name="seg_cls_prob" + ext, type="Softmax", bottom=["seg_cls_score" + ext],
                      top=["seg_cls_prob" + ext]))
        L.append(_ip("bbox_pred" + ext, "join_box_mask" + ext, "bbox_pred" + ext, 84))
        if ext == "":
            L.append(dict(name="stage_bridge", type="Python",
                          bottom=["rois", "bbox_pred", "seg_cls_prob", "im_info"],
                          top=["rois_ext"],
                          python_param=dict(module="pylayer.stage_bridge_layer",
                                            layer="StageBridgeLayer")))
    return L


# ------------------------------------------------------------------ minimal prototxt text reader
_TOKEN = re.compile(r'\s*(?:(#[^\n]*)|([{}])|([A-Za-z_][A-Za-z0-9_]*)\s*:?|"([^"]*)"|\'([^\']*)\'|([^\s{}#]+))')


def parse_prototxt(text):
    """Protobuf text format -> nested dict; repeated keys become lists."""
    pos = 0
    toks = []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            break
        pos = m.end()
        if m.group(1) is not None:
            continue
        if m.group(2):
            toks.append(("brace", m.group(2)))
        elif m.group(3):
            toks.append(("key", m.group(3)))
        elif m.group(4) is not None:
            toks.append(("val", m.group(4)))
        elif m.group(5) is not None:
            toks.append(("val", m.group(5)))
        elif m.group(6):
            toks.append(("val", m.group(6)))

    def conv(v):
        try:
            return int(v)
        except ValueError:
            try:
                return float(v)
            except ValueError:
                return v

    def block(i):
        d = {}
        while i < len(toks):
            kind, v = toks[i]
            if kind == "brace" and v == "}":
                return d, i + 1
            assert kind == "key", "prototxt parse error near %r" % (toks[i],)
            key = v
            nk, nv = toks[i + 1]
            if nk == "brace" and nv == "{":
                val, i = block(i + 2)
            else:
                # a bare word after `key:` lexes as key (enum like MAX) or val
                val, i = conv(nv), i + 2
            if key in d:
                if not isinstance(d[key], list):
                    d[key] = [d[key]]
                d[key].append(val)
            else:
                d[key] = val
        return d, i

    return block(0)[0]


def check_prototxt(path):
    """Raise unless `path` describes the MNC 5-stage test graph (names, types, wiring and the
    parameters this engine bakes in).  Returns the parsed layer list."""
    with open(path) as f:
        net = parse_prototxt(f.read())
    layers = net.get("layer", [])
    if not isinstance(layers, list):
        layers = [layers]
    ref = build_graph()
    if len(layers) != len(ref):
        raise ValueError("unsupported net: %d layers, MNC 5-stage test graph has %d" % (len(layers), len(ref)))
    for got, want in zip(layers, ref):
        as_list = lambda v: v if isinstance(v, list) else ([] if v is None else [v])
        if got.get("name") != want["name"] or got.get("type") != want["type"]:
            raise ValueError("unsupported net: layer %r/%r where %r/%r expected" % (
                got.get("name"), got.get("type"), want["name"], want["type"]))
        if as_list(got.get("bottom")) != want["bottom"] or as_list(got.get("top")) != want["top"]:
            raise ValueError("unsupported net: wiring of layer %r differs" % want["name"])
        for pk in ("convolution_param", "inner_product_param", "roi_warping_param",
                   "mask_resize_param", "pooling_param"):
            if pk in want:
                for k, v in want[pk].items():
                    gv = got.get(pk, {}).get(k)
                    if gv != v and not (isinstance(v, float) and abs(float(gv) - v) < 1e-9):
                        raise ValueError("unsupported net: %s.%s.%s = %r, engine implements %r" % (
                            want["name"], pk, k, gv, v))
    return layers
