"""Minimal ONNX reader for MLP policies.

The reference runs whatever ONNX file the user picks through onnxruntime (`core/policy.py:5-21`).  Neither `onnx` nor
`onnxruntime` is installable here, so this module decodes the protobuf wire format directly and extracts the layer
chain of a feed-forward policy: `Gemm` (any transB / alpha = beta = 1) or `MatMul` + `Add`, separated by one kind of
activation (`Elu`, `Tanh`, `Relu`); `Identity`, `Flatten`, `Reshape`, `Squeeze`, `Unsqueeze` and a trailing `Clip` / `Tanh`-free
output are passed through.  Anything else raises `ValueError` with the offending op.

Field numbers follow onnx.proto3 (ModelProto.graph = 7; GraphProto.node = 1, .initializer = 5; NodeProto.input = 1,
.output = 2, .op_type = 4, .attribute = 5; TensorProto.dims = 1, .data_type = 2, .float_data = 4, .name = 8, .raw_data = 9;
AttributeProto.name = 1, .f = 2, .i = 3).
"""
import struct

import numpy as np


def _varint(buf, i):
    v, shift = 0, 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, i
        shift += 7


def _fields(buf):
    """Yield (field number, wire type, value) of one message; length-delimited values are memoryviews."""
    i, n = 0, len(buf)
    while i < n:
        key, i = _varint(buf, i)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 1:
            v, i = bytes(buf[i:i + 8]), i + 8
        elif wt == 2:
            ln, i = _varint(buf, i)
            v, i = buf[i:i + ln], i + ln
        elif wt == 5:
            v, i = bytes(buf[i:i + 4]), i + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _tensor(buf):
    dims, dtype, name, raw, floats = [], 1, "", None, []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            if wt == 2:
                j = 0
                while j < len(v):
                    d, j = _varint(v, j)
                    dims.append(d)
            else:
                dims.append(v)
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 4:
            floats.append(np.frombuffer(bytes(v), dtype="<f4") if wt == 2 else np.array(struct.unpack("<f", v), dtype=np.float32))
    if dtype == 1:
        arr = np.frombuffer(raw, dtype="<f4") if raw is not None else (np.concatenate(floats) if floats else np.zeros(0, np.float32))
    elif dtype == 11:
        arr = np.frombuffer(raw, dtype="<f8").astype(np.float32)
    elif dtype in (6, 7):
        arr = np.frombuffer(raw, dtype="<i4" if dtype == 6 else "<i8") if raw is not None else np.zeros(0, np.int64)
    else:
        raise ValueError(f"initializer '{name}': unsupported ONNX data type {dtype}")
    return name, np.array(arr).reshape(dims) if dims else np.array(arr)


def _node(buf):
    ins, outs, op, attrs = [], [], "", {}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            ins.append(bytes(v).decode())
        elif fno == 2:
            outs.append(bytes(v).decode())
        elif fno == 4:
            op = bytes(v).decode()
        elif fno == 5:
            an, af, ai = "", None, None
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    an = bytes(v2).decode()
                elif f2 == 2:
                    af = struct.unpack("<f", v2)[0]
                elif f2 == 3:
                    ai = v2
            attrs[an] = af if af is not None else ai
    return op, ins, outs, attrs


def read_graph(path):
    buf = memoryview(open(path, "rb").read())
    graph = None
    for fno, wt, v in _fields(buf):
        if fno == 7:
            graph = v
    if graph is None:
        raise ValueError("not an ONNX model: no graph")
    nodes, inits = [], {}
    for fno, wt, v in _fields(graph):
        if fno == 1:
            nodes.append(_node(v))
        elif fno == 5:
            n, a = _tensor(v)
            inits[n] = a
    return nodes, inits


_PASS = {"Identity", "Flatten", "Reshape", "Squeeze", "Unsqueeze", "Cast", "Dropout", "Transpose", "Concat", "Slice", "Gather", "Shape"}
_ACT = {"Elu": "elu", "Tanh": "tanh", "Relu": "relu"}


def load_policy(path):
    """-> {"kind": "mlp", "layers", "activation"} or {"kind": "lstm", "pre", "lstm": (W[4H,in], R[4H,H], B[8H]), "post", "activation"}.
    An LSTM graph is a chain  [MLP encoder] -> LSTM (one direction, ONNX gate order i, o, f, c) -> [MLP head]."""
    nodes, inits = read_graph(path)
    k = [i for i, n in enumerate(nodes) if n[0] == "LSTM"]
    if not k:
        layers, act = _chain(nodes, inits)
        return {"kind": "mlp", "layers": layers, "activation": act or "elu"}
    if len(k) > 1:
        raise ValueError("more than one LSTM node is not supported")
    op, ins, outs, attrs = nodes[k[0]]
    W, R = inits[ins[1]], inits[ins[2]]
    if W.shape[0] != 1 or attrs.get("direction", 0) not in (0, None):
        raise ValueError("only single-direction LSTMs are supported")
    H = R.shape[2]
    B = inits[ins[3]][0] if len(ins) > 3 and ins[3] in inits else np.zeros(8 * H, np.float32)
    pre, act1 = _chain(nodes[:k[0]], inits, allow_empty=True)
    pre_ops = [n[0] for n in nodes[:k[0]] if n[0] in _ACT or n[0] in ("Gemm", "MatMul", "Add")]
    pre_activated = bool(pre_ops) and pre_ops[-1] in _ACT          # is the encoder's last linear layer followed by the activation?
    post, act2 = _chain(nodes[k[0] + 1:], inits, allow_empty=True)
    return {"kind": "lstm", "pre": pre, "post": post, "lstm": (np.ascontiguousarray(W[0], np.float32), np.ascontiguousarray(R[0], np.float32),
            np.ascontiguousarray(B, np.float32)), "activation": act1 or act2 or "elu", "pre_activated": pre_activated}


def load_mlp(path):
    nodes, inits = read_graph(path)
    layers, act = _chain(nodes, inits)
    return layers, act or "elu"


def _chain(nodes, inits, allow_empty=False):
    """-> (layers [(W[out, in], b[out]), ...], activation name or None).  The output clip to [-1, 1] is applied by the policy
    kernel (core/policy.py:20), so a trailing Clip node is accepted and dropped."""
    layers, act, pending = [], None, None
    for op, ins, outs, attrs in nodes:
        if op == "Gemm":
            if attrs.get("alpha", 1.0) != 1.0 or attrs.get("beta", 1.0) != 1.0 or attrs.get("transA", 0):
                raise ValueError("Gemm with alpha/beta != 1 or transA is not supported")
            w = inits[ins[1]]
            w = w if attrs.get("transB", 0) else w.T
            b = inits[ins[2]] if len(ins) > 2 else np.zeros(w.shape[0], np.float32)
            layers.append((np.ascontiguousarray(w, np.float32), np.ascontiguousarray(b, np.float32).reshape(-1)))
        elif op == "MatMul":
            wname = ins[1] if ins[1] in inits else ins[0]
            pending = np.ascontiguousarray(inits[wname].T, np.float32)
            layers.append((pending, np.zeros(pending.shape[0], np.float32)))
        elif op == "Add" and pending is not None and any(i in inits for i in ins):
            b = inits[[i for i in ins if i in inits][0]]
            layers[-1] = (layers[-1][0], np.ascontiguousarray(b, np.float32).reshape(-1))
            pending = None
        elif op in _ACT:
            if act is not None and act != _ACT[op]:
                raise ValueError(f"mixed hidden activations ({act}, {_ACT[op]}) are not supported")
            act = _ACT[op]
        elif op in _PASS or op == "Clip" or op == "Constant":
            continue
        else:
            raise ValueError(f"unsupported ONNX op '{op}' in an MLP policy")
    if not layers and not allow_empty:
        raise ValueError("no Gemm / MatMul layers found")
    for (w0, _), (w1, _) in zip(layers[:-1], layers[1:]):
        if w1.shape[1] != w0.shape[0]:
            raise ValueError("layer shapes do not chain")
    return layers, act
