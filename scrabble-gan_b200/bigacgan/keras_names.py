"""Which checkpoint key holds which weight: the names Keras gives the reference's variables when it saves a model.

`model.save_weights(prefix)` in TF-checkpoint format (reference data_utils.py:346-348) stores every variable under
    layer_with_weights-<N>/<attribute>/.ATTRIBUTES/VARIABLE_VALUE
where N counts the layers WITH WEIGHTS in the order of `model.layers`, and <attribute> is the Python attribute the layer
keeps the variable in (Conv2D / Conv2DTranspose / Dense: kernel, bias; BatchNormalization: gamma, beta, moving_mean,
moving_variance; the reference's SpatialEmbedding: kernel; its NonLocalBlock: sigma -- the four 1x1 kernels of that block are
re-created in every call and never tracked).

`model.layers` of a functional model is NOT the order of construction: Keras sorts the layers by their depth in the graph
(longest path to an output, deepest first) and breaks ties by the order a depth-first walk from the outputs first meets them
(tf.keras Network._map_graph_network, TF 2.1).  Every raw TensorFlow op applied to a Keras tensor (tf.nn.relu, `net *= gamma`,
tf.reshape with a dynamic shape, tf.split, ...) is a layer of its own in that graph (TensorFlowOpLayer), so the op structure
matters for the depths.  This module therefore re-states the reference's model graphs (net_architecture.py:9-79, 182-296,
299-355; resnet_ops.py:13-28, 46-74, 93-115) op by op as a tiny symbolic graph and runs Keras' own ordering rule on it.

PARITY UNPINNED: neither TensorFlow nor a reference checkpoint is available here.  The rule and the graphs are re-stated from
source; `load_keras_checkpoint` therefore validates EVERY tensor's shape against the model, reports the full key table on
any mismatch, and accepts an explicit `key_map` to override the derived names."""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


# ----------------------------------------------------------------------------------------------------
# a symbolic stand-in for the Keras functional graph
# ----------------------------------------------------------------------------------------------------
class _Layer:
    def __init__(self, name: str, weights: Sequence[Tuple[str, str]] = ()):
        self.name = name
        self.weights = list(weights)          # (keras attribute, libsgan variable name)
        self.inbound: List["_Layer"] = []

    def __call__(self, *inputs: "_Layer") -> "_Layer":
        assert not self.inbound, "every layer of these graphs is called once"
        self.inbound = list(inputs)
        return self


def _op(name: str, *inputs: _Layer) -> _Layer:
    """A raw TF op on Keras tensors = one TensorFlowOpLayer whose inbound tensors are the op's tensor inputs, in order."""
    return _Layer("tf_op_layer_" + name)(*inputs)


def keras_layer_order(outputs: Sequence[_Layer]) -> List[_Layer]:
    """Network._map_graph_network (TF 2.1): traversal indices are given when a layer is first reached by the depth-first
    walk from the outputs (before its inputs are walked); a node's depth is the longest path to an output; layers are
    listed deepest first, ties in traversal order."""
    layer_index: Dict[_Layer, int] = {}
    finished, post_order = set(), []

    def build_map(layer: _Layer):
        stack = [(layer, 0)]
        while stack:                                   # iterative DFS (the graphs are a few hundred nodes deep)
            lay, i = stack.pop()
            if i == 0:
                if lay in finished:
                    continue
                if lay not in layer_index:
                    layer_index[lay] = len(layer_index)
            if i < len(lay.inbound):
                stack.append((lay, i + 1))
                nxt = lay.inbound[i]
                if nxt not in finished:
                    stack.append((nxt, 0))
            else:
                if lay not in finished:
                    finished.add(lay)
                    post_order.append(lay)

    for out in outputs:
        build_map(out)
    depth: Dict[_Layer, int] = {}
    for lay in reversed(post_order):
        d = depth.setdefault(lay, 0)
        for src in lay.inbound:
            depth[src] = max(depth.get(src, 0), d + 1)
    return sorted(post_order, key=lambda lay: (-depth[lay], layer_index[lay]))


def _weight_keys(outputs: Sequence[_Layer]) -> "OrderedDict[str, str]":
    """{libsgan variable name: checkpoint key} of a model graph."""
    out: "OrderedDict[str, str]" = OrderedDict()
    n = 0
    for lay in keras_layer_order(outputs):
        if not lay.weights:
            continue
        for attr, ours in lay.weights:
            out[ours] = "layer_with_weights-{}/{}{}".format(n, attr, SUFFIX)
        n += 1
    return out


# ----------------------------------------------------------------------------------------------------
# the reference's building blocks, op by op
# ----------------------------------------------------------------------------------------------------
def _conv(pre: str, bias: bool = True) -> _Layer:
    return _Layer("conv2d:" + pre, [("kernel", pre + ".w")] + ([("bias", pre + ".b")] if bias else []))


def _res_block_down(x: _Layer, pre: str, is_last: bool) -> _Layer:
    """resnet_ops.py:93-115."""
    net = _op("Relu", x)
    net = _conv(pre + ".conv1")(net)
    net = _op("Relu", net)
    net = _conv(pre + ".conv2")(net)
    if not is_last:
        net = _op("AvgPool", net)
    short = _conv(pre + ".short")(x)
    if not is_last:
        short = _op("AvgPool", short)
    return _op("AddV2", net, short)


def _non_local(x: _Layer, pre: str) -> _Layer:
    """arch_ops.py:32-67 as ONE layer (it is invoked through __call__); only `sigma` is a tracked weight."""
    return _Layer("NonLocalBlock:" + pre, [("sigma", pre + ".sigma")])(x)


def _cbn(x: _Layer, z_i: _Layer, pre: str) -> _Layer:
    """resnet_ops.py:13-28: BN(scale=False, center=False); gamma = Dense(z); reshape; mul; beta = Dense(z); reshape; add."""
    net = _Layer("bn:" + pre, [("moving_mean", pre + ".moving_mean"), ("moving_variance", pre + ".moving_var")])(x)
    gamma = _Layer("dense:" + pre + ".gamma", [("kernel", pre + ".gamma.w")])(z_i)
    gamma = _op("Reshape", gamma)
    net = _op("Mul", net, gamma)
    beta = _Layer("dense:" + pre + ".beta", [("kernel", pre + ".beta.w")])(z_i)
    beta = _op("Reshape", beta)
    return _op("AddV2", net, beta)


def _res_block_up(x: _Layer, z_i: _Layer, pre: str) -> _Layer:
    """resnet_ops.py:46-74."""
    net = _cbn(x, z_i, pre + ".cbn1")
    net = _op("Relu", net)
    net = _Layer("conv2d_transpose:" + pre + ".up", [("kernel", pre + ".up.w"), ("bias", pre + ".up.b")])(net)
    net = _cbn(net, z_i, pre + ".cbn2")
    net = _op("Relu", net)
    net = _conv(pre + ".conv")(net)
    short = _Layer("conv2d_transpose:" + pre + ".short", [("kernel", pre + ".short.w"), ("bias", pre + ".short.b")])(x)
    return _op("AddV2", net, short)


def _down_trunk(x: _Layer, prefix: str, attn_after: Callable[[str], bool]) -> _Layer:
    net = x
    for i in range(4):
        name = "{}{}".format(prefix, i + 1)
        net = _res_block_down(net, name, i == 3)
        if attn_after(name):
            net = _non_local(net, name + ".attn")
    net = _op("Relu", net)
    return _Layer("global_average_pooling2d")(net)


# ----------------------------------------------------------------------------------------------------
# the reference's models
# ----------------------------------------------------------------------------------------------------
def recognizer_keys() -> "OrderedDict[str, str]":
    """make_recognizer (net_architecture.py:9-79): a chain, so the order is the order of construction."""
    x = _Layer("input_images")
    net = x
    plan = [("conv1", None), ("pool", None), ("conv2", None), ("pool", None), ("conv3", None), ("conv4", None), ("pool", None),
            ("conv5", "bn5"), ("conv6", "bn6"), ("pool", None), ("conv7", None)]
    for name, bn in plan:
        if name == "pool":
            net = _Layer("max_pooling2d")(net)
            continue
        net = _conv(name)(net)
        if bn:
            net = _Layer("bn:" + bn, [("gamma", bn + ".gamma"), ("beta", bn + ".beta"), ("moving_mean", bn + ".moving_mean"),
                                      ("moving_variance", bn + ".moving_var")])(net)
    net = _Layer("lambda_squeeze")(net)
    net = _Layer("dense", [("kernel", "dense.w"), ("bias", "dense.b")])(net)
    labels, in_len, lab_len = _Layer("y_true"), _Layer("input_length"), _Layer("label_length")
    out = _Layer("ctc")(labels, net, in_len, lab_len)
    return _weight_keys([out])


def discriminator_keys(blocks_with_attention: str = "B1") -> "OrderedDict[str, str]":
    """make_discriminator / make_style_promoter (net_architecture.py:299-355, 358-414)."""
    x = _Layer("input")
    feats = _down_trunk(x, "B", lambda nm: nm in blocks_with_attention)
    out = _Layer("dense", [("kernel", "dense.w")])(feats)
    return _weight_keys([out])


def generator_keys(blocks_with_attention: str = "B3", style_encoder: bool = True) -> "OrderedDict[str, str]":
    """make_generator of this fork (net_architecture.py:182-296: style encoder -> z) or, with style_encoder=False, the
    upstream form whose first input is z itself (the signature run_inference.py:35 feeds)."""
    y = _Layer("input_labels")
    se = _Layer("SpatialEmbedding", [("kernel", "filter_bank")])(y)
    if style_encoder:
        imgs = _Layer("input_images")
        feats = _down_trunk(imgs, "B_style", lambda nm: nm == "B_style1")
        z = _Layer("dense:style", [("kernel", "style_dense.w")])(feats)
    else:
        z = _Layer("input_z")
    split = _op("split", z)                                        # one op, four outputs: z0 and z_per_block
    z0 = _op("Reshape", split)
    # tf.tile(z0, [1, tf.shape(se_layer)[1], 1, 1]): Shape -> strided_slice -> Pack (multiples) -> Tile
    mult = _op("Pack", _op("strided_slice", _op("Shape", se)))
    net = _op("BatchMatMulV2", _op("Tile", z0, mult), se)
    net = _op("Squeeze", net)
    # two reshapes with tf.shape(net)[0] in the target shape, then the transpose (net_architecture.py:269-271)
    net = _op("Reshape", net, _op("Pack", _op("strided_slice", _op("Shape", net))))
    net = _op("Reshape", net, _op("Pack", _op("strided_slice", _op("Shape", net))))
    net = _op("Transpose", net)
    for i in range(3):
        name = "B{}".format(i + 1)
        net = _res_block_up(net, split, name)
        if name in blocks_with_attention:
            net = _non_local(net, name + ".attn")
    net = _Layer("bn:final", [("gamma", "bn.gamma"), ("beta", "bn.beta"), ("moving_mean", "bn.moving_mean"),
                              ("moving_variance", "bn.moving_var")])(net)
    net = _op("Relu", net)
    net = _conv("out")(net)
    out = _op("Tanh", net)
    return _weight_keys([out])


# ----------------------------------------------------------------------------------------------------
# loading / saving libsgan models in the reference's checkpoint format
# ----------------------------------------------------------------------------------------------------
def keys_for(model) -> "OrderedDict[str, str]":
    kind = type(model).__name__
    if kind == "Recognizer":
        return recognizer_keys()
    if kind == "Discriminator":
        return discriminator_keys("".join(sorted("B%d" % (i + 1) for i in model.trunk.attn)))
    if kind == "Generator":
        return generator_keys("".join(sorted("B%d" % (i + 1) for i in model.attn)), style_encoder=model.style is not None)
    raise TypeError("no Keras naming for {}".format(kind))


def load_keras_checkpoint(model, prefix: str, key_map: Optional[Dict[str, str]] = None, strict: bool = True) -> List[str]:
    """Load `<prefix>.index/.data-*` written by the reference's `model.save_weights(prefix)` into a libsgan model.  Every
    variable the checkpoint can provide is shape-checked; NonLocalBlock's four 1x1 kernels are not in a reference checkpoint
     and keep their current values.  Returns the names of the variables that were loaded."""
    from .. import tf_checkpoint
    tensors = tf_checkpoint.read_checkpoint(prefix)
    keys = dict(keys_for(model))
    keys.update(key_map or {})
    state, problems = {}, []
    for v in model.store.vars:
        key = keys.get(v.name)
        if key is None:
            continue                                    # untracked in the reference (attention projections)
        if key not in tensors:
            problems.append("{}: checkpoint has no key {}".format(v.name, key))
            continue
        a = tensors[key]
        if v.name.endswith(".sigma"):
            a = a.reshape(v.shape)
        if tuple(a.shape) != tuple(v.shape):
            problems.append("{}: checkpoint key {} has shape {}, the model needs {}".format(v.name, key, tuple(a.shape), tuple(v.shape)))
            continue
        state[v.name] = a
    if problems and strict:
        table = "\\n".join("  {:<60s} {}".format(k, tuple(t.shape)) for k, t in tensors.items())
        raise ValueError("checkpoint {} does not match the derived Keras names (pass key_map={{variable: key}} to override):\\n  {}\\n"
                         "checkpoint contents:\\n{}".format(prefix, "\\n  ".join(problems), table))
    model.store.load_state_dict(state, strict=False)
    return sorted(state)


def save_keras_checkpoint(model, prefix: str) -> None:
    """Write the model as the reference's `save_weights(prefix)` would name it (TF-checkpoint format, Keras keys)."""
    from .. import tf_checkpoint
    keys = keys_for(model)
    sd = model.store.state_dict()
    tensors = {}
    for name, key in keys.items():
        a = sd[name].detach().cpu().numpy().astype(np.float32)
        tensors[key] = a.reshape(()) if name.endswith(".sigma") else a
    tf_checkpoint.write_checkpoint(prefix, tensors)
