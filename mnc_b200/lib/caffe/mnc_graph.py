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


def _trunk():
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
    return L


def build_frcnn_graph():
    """models/VGG16/faster_rcnn_end2end/test.prototxt (the RPN layers are named `rpn_conv/3x3`,
    `rpn/output` there; Dropout is the identity in TEST phase)."""
    L = _trunk()
    L.append(_conv("rpn_conv/3x3", "conv5_3", "rpn/output", 512))
    L.append(_relu("rpn_relu/3x3", "rpn/output"))
    L.append(_conv("rpn_cls_score", "rpn/output", "rpn_cls_score", 18, k=1, pad=0))
    L.append(_conv("rpn_bbox_pred", "rpn/output", "rpn_bbox_pred", 36, k=1, pad=0))
    L.append(dict(name="rpn_cls_score_reshape", type="Reshape", bottom=["rpn_cls_score"],
                  top=["rpn_cls_score_reshape"]))
    L.append(dict(name="rpn_cls_prob", type="Softmax", bottom=["rpn_cls_score_reshape"],
                  top=["rpn_cls_prob"]))
    L.append(dict(name="rpn_cls_prob_reshape", type="Reshape", bottom=["rpn_cls_prob"],
                  top=["rpn_cls_prob_reshape"]))
    L.append(dict(name="proposal", type="Python",
                  bottom=["rpn_cls_prob_reshape", "rpn_bbox_pred", "im_info"], top=["rois"],
                  python_param=dict(module="pylayer.proposal_layer", layer="ProposalLayer")))
    L.append(dict(name="roi_pool5", type="ROIWarping", bottom=["conv5_3", "rois"], top=["pool5"],
                  roi_warping_param=dict(pooled_w=7, pooled_h=7, spatial_scale=0.0625)))
    L.append(_ip("fc6", "pool5", "fc6", 4096))
    L.append(_relu("relu6", "fc6"))
    L.append(dict(name="drop6", type="Dropout", bottom=["fc6"], top=["fc6"]))
    L.append(_ip("fc7", "fc6", "fc7", 4096))
    L.append(_relu("relu7", "fc7"))
    L.append(dict(name="drop7", type="Dropout", bottom=["fc7"], top=["fc7"]))
    L.append(_ip("cls_score", "fc7", "cls_score", 21))
    L.append(dict(name="cls_prob", type="Softmax", bottom=["cls_score"], top=["cls_prob"]))
    L.append(_ip("bbox_pred", "fc7", "bbox_pred", 84))
    return L


def build_cfm_graph():
    """models/VGG16/cfm/test.prototxt: inputs data, rois, masks; no RPN."""
    L = _trunk()
    rp = lambda name, top, p: dict(name=name, type="ROIPooling", bottom=["conv5_3", "rois"], top=[top],
                                   roi_pooling_param=dict(pooled_w=p, pooled_h=p, spatial_scale=0.0625))
    L.append(rp("roi_pooling_conv5", "roi_pooling_conv5", 7))
    L.append(_ip("fc6", "roi_pooling_conv5", "fc6", 4096))
    L.append(_relu("relu6", "fc6"))
    L.append(_ip("fc7", "fc6", "fc7", 4096))
    L.append(_relu("relu7", "fc7"))
    L.append(rp("roi_pooling_conv5_mask", "roi_pooling_conv5_mask", 14))
    L.append(dict(name="mask_pooling", type="MaskPooling", bottom=["roi_pooling_conv5_mask", "masks"],
                  top=["roi_mask_conv5"]))
    L.append(_pool("roi_mask_conv5", "roi_mask_conv5", "roi_mask_conv5_pool"))
    L.append(_ip("fc6_mask", "roi_mask_conv5_pool", "fc6_mask", 4096))
    L.append(_relu("relu6_mask", "fc6_mask"))
    L.append(_ip("fc7_mask", "fc6_mask", "fc7_mask", 4096))
    L.append(_relu("relu7_mask", "fc7_mask"))
    L.append(_ip("fc6_maskest", "roi_pooling_conv5_mask", "fc6_maskest", 256))
    L.append(_relu("relu6_maskest", "fc6_maskest"))
    L.append(_ip("mask_pred", "fc6_maskest", "mask_pred", 441))
    L.append(dict(name="mask_prob", type="Sigmoid", bottom=["mask_pred"], top=["mask_prob"]))
    L.append(dict(name="join_box_mask", type="Concat", bottom=["fc7_mask", "fc7"], top=["join_box_mask"]))
    L.append(_ip("cls_score", "join_box_mask", "cls_score", 21))
    L.append(dict(name="cls_prob", type="Softmax", bottom=["cls_score"], top=["cls_prob"]))
    L.append(_ip("seg_cls_score", "join_box_mask", "seg_cls_score", 21))
    L.append(dict(name="seg_cls_prob", type="Softmax", bottom=["seg_cls_score"], top=["seg_cls_prob"]))
    L.append(_ip("bbox_pred", "join_box_mask", "bbox_pred", 84))
    return L


GRAPHS = {"mnc_5stage": lambda: build_graph(), "faster_rcnn": build_frcnn_graph, "cfm": build_cfm_graph}
GRAPH_INPUTS = {"mnc_5stage": ["data", "im_info"], "faster_rcnn": ["data", "im_info"],
                "cfm": ["data", "rois", "masks"]}


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


def _as_list(v):
    return v if isinstance(v, list) else ([] if v is None else [v])


def _mismatch(layers, ref):
    """None if the parsed layer list equals the reference graph, else a description."""
    if len(layers) != len(ref):
        return "%d layers where %d expected" % (len(layers), len(ref))
    for got, want in zip(layers, ref):
        if got.get("name") != want["name"] or got.get("type") != want["type"]:
            return "layer %r/%r where %r/%r expected" % (got.get("name"), got.get("type"),
                                                        want["name"], want["type"])
        if _as_list(got.get("bottom")) != want["bottom"] or _as_list(got.get("top")) != want["top"]:
            return "wiring of layer %r differs" % want["name"]
        for pk in ("convolution_param", "inner_product_param", "roi_warping_param",
                   "roi_pooling_param", "mask_resize_param", "pooling_param"):
            if pk in want:
                for k, v in want[pk].items():
                    gv = got.get(pk, {}).get(k)
                    if gv != v and not (isinstance(v, float) and gv is not None and abs(float(gv) - v) < 1e-9):
                        return "%s.%s.%s = %r, engine implements %r" % (want["name"], pk, k, gv, v)
    return None


def identify_prototxt(path):
    """Which of the supported test graphs `path` describes: "mnc_5stage" (models/VGG16/mnc_5stage/
    test.prototxt), "faster_rcnn" (faster_rcnn_end2end/test.prototxt) or "cfm" (cfm/test.prototxt);
    names, types, wiring and the parameters the engines bake in must all match.  -> (kind, layers)."""
    with open(path) as f:
        net = parse_prototxt(f.read())
    layers = _as_list(net.get("layer", []))
    why = {}
    for kind, build in GRAPHS.items():
        why[kind] = _mismatch(layers, build())
        if why[kind] is None:
            if _as_list(net.get("input")) != GRAPH_INPUTS[kind]:
                raise ValueError("unsupported net: inputs %r, %s takes %r" % (
                    net.get("input"), kind, GRAPH_INPUTS[kind]))
            return kind, layers
    raise ValueError("unsupported net: " + "; ".join("not %s (%s)" % kv for kv in why.items()))


def check_prototxt(path, kind="mnc_5stage"):
    """Raise unless `path` describes the given test graph.  Returns the parsed layer list."""
    got, layers = identify_prototxt(path)
    if got != kind:
        raise ValueError("unsupported net: %s is the %s graph, %s expected" % (path, got, kind))
    return layers
