"""Batched policy: drop-in for core/policy.py (MLPPolicy.get_action, /root/reference/core/policy.py:11-21).

`MLPPolicy.get_action(state[N, state_dim]) -> action[N, action_dim]`, clipped to [-1, 1], computed by the
hand-written tcgen05/TMEM kernel in csrc/policy.cu (bf16 operands, fp32 accumulate).  The reference loads
an ONNX file chosen in the GUI; no weights ship with it (`weights/tmp` is empty), so `synthetic_mlp`
builds the stand-in network of SURVEY.md section 8d (state_dim -> 512 -> 256 -> 128 -> action_dim, ELU).
"""
import ctypes

import numpy as np
import torch

from . import lib as _libmod

ACTIVATIONS = {"elu": 0, "tanh": 1, "relu": 2}


def synthetic_mlp(state_dim, action_dim, hidden=(512, 256, 128), seed=1234):
    """Weights ~ N(0, 1/fan_in), small biases; deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    dims = [int(state_dim)] + [int(h) for h in hidden] + [int(action_dim)]
    layers = []
    for i in range(len(dims) - 1):
        w = (rng.standard_normal((dims[i + 1], dims[i])) / np.sqrt(dims[i])).astype(np.float32)
        b = (0.01 * rng.standard_normal(dims[i + 1])).astype(np.float32)
        layers.append((w, b))
    return layers


class MLPPolicy:
    def __init__(self, layers, activation="elu", device="cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("cosim_b200 policy needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device)
        self.layers = [(np.ascontiguousarray(w, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)) for w, b in layers]
        self.activation = activation
        dims = [self.layers[0][0].shape[1]] + [w.shape[0] for w, _ in self.layers]
        for (w, b), din, dout in zip(self.layers, dims[:-1], dims[1:]):
            assert w.shape == (dout, din) and b.shape == (dout,), "layers must be (W[out, in], b[out]) pairs that chain"
        self.state_dim, self.action_dim = dims[0], dims[-1]
        self._L = _libmod.lib()
        n = len(self.layers)
        d = (ctypes.c_int * (n + 1))(*dims)
        wp = (ctypes.c_void_p * n)(*[w.ctypes.data for w, _ in self.layers])
        bp = (ctypes.c_void_p * n)(*[b.ctypes.data for _, b in self.layers])
        self._h = ctypes.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = self._L.cosim_policy_create(idx, n, d, wp, bp, ACTIVATIONS[activation], ctypes.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"cosim_policy_create failed with code {rc}")
        self._action = None
        self._state_dev = None

    def get_action(self, state):
        """state: float32 CUDA tensor [N, state_dim] -> float32 CUDA tensor [N, action_dim] in [-1, 1]."""
        state = torch.as_tensor(state, dtype=torch.float32, device=self.device)
        squeeze = state.dim() == 1
        if squeeze:
            state = state.unsqueeze(0)
        state = state.contiguous()
        n = state.shape[0]
        if state.shape[1] != self.state_dim:
            raise RuntimeError(f"The current state length (={state.shape[1]}) does not match the input length expected by the policy (={self.state_dim}).")
        if self._action is None or self._action.shape[0] != n:
            self._action = torch.empty((n, self.action_dim), dtype=torch.float32, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self._L.cosim_policy_forward(self._h, ctypes.c_void_p(state.data_ptr()), n, ctypes.c_void_p(self._action.data_ptr()), stream)
        if rc != 0:
            raise RuntimeError(f"cosim_policy_forward failed with code {rc}")
        return self._action[0] if squeeze else self._action

    def get_action_host(self, state_host, action_host):
        """numpy (pinned) in / out: H2D state, forward, D2H action, sync -- the reference's call shape."""
        n = state_host.shape[0]
        if self._state_dev is None or self._state_dev.shape[0] != n:
            self._state_dev = torch.empty((n, self.state_dim), dtype=torch.float32, device=self.device)
        self._state_dev.copy_(torch.from_numpy(state_host), non_blocking=True)
        a = self.get_action(self._state_dev)
        torch.from_numpy(action_host).copy_(a, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    @property
    def launch_count(self):
        return self._L.cosim_policy_launch_count(self._h)

    def reference_forward(self, state, emulate_bf16=True):
        """Plain PyTorch fp32 reference of the same op (tests): optional bf16 rounding of operands."""
        x = torch.as_tensor(state, dtype=torch.float32, device=self.device)
        f = {"elu": torch.nn.functional.elu, "tanh": torch.tanh, "relu": torch.relu}[self.activation]
        rnd = (lambda t: t.to(torch.bfloat16).float()) if emulate_bf16 else (lambda t: t)
        for i, (w, b) in enumerate(self.layers):
            wt = rnd(torch.from_numpy(w).to(self.device))
            x = rnd(x) @ wt.t() + torch.from_numpy(b).to(self.device)
            if i < len(self.layers) - 1:
                x = f(x)
        return x.clamp(-1, 1)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cosim_policy_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_policy(config, policy_path=None, state_dim=None, action_dim=None, device="cuda:0"):
    """core/policy.py:49-53.  MLP policies are read from the user's ONNX file (cosim_b200/onnx_reader.py); without a file the
    synthetic MLP of SURVEY.md section 8d is used.  LSTM policies are a 'next' row."""
    if config.get("policy", {}).get("use_lstm"):
        raise NotImplementedError("LSTM policies are not implemented yet (SURVEY.md section 8f, row 1)")
    if policy_path:
        from .onnx_reader import load_mlp
        layers, act = load_mlp(policy_path)          # the user's ONNX file, as in core/policy.py:7-9
        return MLPPolicy(layers, act, device)
    return MLPPolicy(synthetic_mlp(state_dim, action_dim), "elu", device)
