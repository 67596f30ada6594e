"""The ONNX reader decodes protobuf directly (onnx / onnxruntime are not installable here): build a small policy file by
hand-encoding the wire format and read it back."""
import struct

import numpy as np
import pytest

from cosim_b200.onnx_reader import load_mlp


def _vi(v):
    out = b""
    while True:
        b = v & 0x7F
        v >>= 7
        out += bytes([b | (0x80 if v else 0)])
        if not v:
            return out


def _ld(fno, payload):
    return _vi((fno << 3) | 2) + _vi(len(payload)) + payload


def _tensor(name, arr, raw=True):
    arr = np.asarray(arr, np.float32)
    msg = b"".join(_vi((1 << 3) | 0) + _vi(d) for d in arr.shape) + _vi((2 << 3) | 0) + _vi(1) + _ld(8, name.encode())
    msg += _ld(9, arr.tobytes()) if raw else _ld(4, arr.tobytes())
    return msg


def _node(op, ins, outs, attrs=None):
    msg = b"".join(_ld(1, i.encode()) for i in ins) + b"".join(_ld(2, o.encode()) for o in outs) + _ld(4, op.encode())
    for k, v in (attrs or {}).items():
        a = _ld(1, k.encode()) + (_vi((3 << 3) | 0) + _vi(v) if isinstance(v, int) else _vi((2 << 3) | 5) + struct.pack("<f", v))
        msg += _ld(5, a)
    return msg


def _model(nodes, inits):
    graph = b"".join(_ld(1, n) for n in nodes) + _ld(2, b"policy") + b"".join(_ld(5, t) for t in inits)
    return _vi((1 << 3) | 0) + _vi(8) + _ld(7, graph)


def test_gemm_and_matmul_chain(tmp_path):
    rng = np.random.default_rng(0)
    w1, b1 = rng.standard_normal((16, 10)).astype(np.float32), rng.standard_normal(16).astype(np.float32)
    w2, b2 = rng.standard_normal((16, 8)).astype(np.float32), rng.standard_normal(8).astype(np.float32)      # [in, out] for MatMul
    w3, b3 = rng.standard_normal((8, 4)).astype(np.float32), rng.standard_normal(4).astype(np.float32)       # Gemm without transB
    nodes = [_node("Gemm", ["x", "w1", "b1"], ["h1"], {"transB": 1, "alpha": 1.0}), _node("Elu", ["h1"], ["a1"]),
             _node("MatMul", ["a1", "w2"], ["m2"]), _node("Add", ["m2", "b2"], ["h2"]), _node("Elu", ["h2"], ["a2"]),
             _node("Gemm", ["a2", "w3", "b3"], ["y"]), _node("Clip", ["y"], ["out"])]
    inits = [_tensor("w1", w1), _tensor("b1", b1, raw=False), _tensor("w2", w2), _tensor("b2", b2), _tensor("w3", w3), _tensor("b3", b3)]
    path = tmp_path / "policy.onnx"
    path.write_bytes(_model(nodes, inits))
    layers, act = load_mlp(str(path))
    assert act == "elu" and [l[0].shape for l in layers] == [(16, 10), (8, 16), (4, 8)]
    np.testing.assert_array_equal(layers[0][0], w1); np.testing.assert_array_equal(layers[0][1], b1)
    np.testing.assert_array_equal(layers[1][0], w2.T); np.testing.assert_array_equal(layers[1][1], b2)
    np.testing.assert_array_equal(layers[2][0], w3.T); np.testing.assert_array_equal(layers[2][1], b3)


def test_rejects_unsupported_ops(tmp_path):
    path = tmp_path / "p.onnx"
    path.write_bytes(_model([_node("LSTM", ["x"], ["y"])], []))
    with pytest.raises(ValueError):
        load_mlp(str(path))


def test_lstm_graph(tmp_path):
    from cosim_b200.onnx_reader import load_policy
    rng = np.random.default_rng(1)
    H, nin = 8, 6
    we, be = rng.standard_normal((nin, 5)).astype(np.float32), rng.standard_normal(nin).astype(np.float32)
    W, R, B = rng.standard_normal((1, 4 * H, nin)).astype(np.float32), rng.standard_normal((1, 4 * H, H)).astype(np.float32), rng.standard_normal((1, 8 * H)).astype(np.float32)
    wh, bh = rng.standard_normal((3, H)).astype(np.float32), rng.standard_normal(3).astype(np.float32)
    nodes = [_node("Gemm", ["state", "we", "be"], ["e"], {"transB": 1}), _node("Tanh", ["e"], ["ea"]), _node("Unsqueeze", ["ea"], ["x"]),
             _node("LSTM", ["x", "W", "R", "B", "", "h_in", "c_in"], ["y", "h_out", "c_out"], {"hidden_size": H}),
             _node("Squeeze", ["y"], ["ys"]), _node("Gemm", ["ys", "wh", "bh"], ["action"], {"transB": 1})]
    inits = [_tensor("we", we), _tensor("be", be), _tensor("W", W), _tensor("R", R), _tensor("B", B), _tensor("wh", wh), _tensor("bh", bh)]
    path = tmp_path / "lstm.onnx"
    path.write_bytes(_model(nodes, inits))
    spec = load_policy(str(path))
    assert spec["kind"] == "lstm" and spec["activation"] == "tanh" and spec["pre_activated"]
    np.testing.assert_array_equal(spec["pre"][0][0], we); np.testing.assert_array_equal(spec["post"][0][0], wh)
    np.testing.assert_array_equal(spec["lstm"][0], W[0]); np.testing.assert_array_equal(spec["lstm"][1], R[0]); np.testing.assert_array_equal(spec["lstm"][2], B[0])
