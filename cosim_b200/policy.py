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
    def __init__(self, layers, activation="elu", device="cuda:0", raw_output=False, activated_output=False):
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
        self.raw_output = bool(raw_output)     # plain linear output of the last layer, no clip (LSTM gates / encoders)
        self.activated_output = bool(activated_output)   # hidden activation on the output instead of the clip (encoders)
        flags = 0x200 if activated_output else (0x100 if raw_output else 0)
        rc = self._L.cosim_policy_create(idx, n, d, wp, bp, ACTIVATIONS[activation] | flags, ctypes.byref(self._h))
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
        if self.activated_output:
            return f(x)
        return x if self.raw_output else x.clamp(-1, 1)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cosim_policy_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LSTMPolicy:
    """Batched drop-in for core/policy.py:24-47 (LSTMPolicy): action = clip(head(LSTM(encoder(state), h, c))), h and c carried
    from step to step per environment (zeros at start, as in the reference).  Every matrix product runs in the tcgen05
    kernel: encoder (optional MLP), gate pre-activations [x | h] -> 4H (ONNX gate order i, o, f, c), head MLP; the cell
    update is the elementwise kernel k_lstm_cell."""

    def __init__(self, lstm, pre_layers=(), post_layers=(), activation="elu", device="cuda:0", num_envs=1, pre_activated=True):
        W, R, B = (np.asarray(a, np.float32) for a in lstm)          # W [4H, in], R [4H, H], B [8H] = [Wb | Rb]
        self.H = R.shape[1]
        assert W.shape[0] == 4 * self.H and R.shape[0] == 4 * self.H and B.shape == (8 * self.H,)
        self.device = torch.device(device)
        self.pre = MLPPolicy(list(pre_layers), activation, device, raw_output=not pre_activated, activated_output=pre_activated) if len(pre_layers) else None
        self.gates = MLPPolicy([(np.concatenate([W, R], axis=1), B[:4 * self.H] + B[4 * self.H:])], activation, device, raw_output=True)
        self.post = MLPPolicy(list(post_layers), activation, device) if len(post_layers) else None
        self.lstm = (W, R, B)
        self.state_dim = self.pre.state_dim if self.pre else W.shape[1]
        self.action_dim = self.post.action_dim if self.post else self.H
        self._L = _libmod.lib()
        self.reset(num_envs=num_envs)

    def reset(self, mask=None, num_envs=None):
        """Zero h / c (all environments, or the rows selected by `mask`)."""
        if num_envs is not None:
            self.h = torch.zeros((num_envs, self.H), dtype=torch.float32, device=self.device)
            self.c = torch.zeros_like(self.h)
        elif mask is None:
            self.h.zero_(); self.c.zero_()
        else:
            m = torch.as_tensor(mask, device=self.device).bool()
            self.h[m] = 0; self.c[m] = 0

    def get_action(self, state):
        state = torch.as_tensor(state, dtype=torch.float32, device=self.device)
        squeeze = state.dim() == 1
        if squeeze:
            state = state.unsqueeze(0)
        n = state.shape[0]
        if self.h.shape[0] != n:
            self.reset(num_envs=n)
        x = self.pre.get_action(state) if self.pre else state
        g = self.gates.get_action(torch.cat([x, self.h], dim=1))
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self._L.cosim_lstm_cell(ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(self.c.data_ptr()), ctypes.c_void_p(self.h.data_ptr()), n, self.H, stream)
        if rc != 0:
            raise RuntimeError(f"cosim_lstm_cell failed with code {rc}")
        a = self.post.get_action(self.h) if self.post else self.h.clamp(-1, 1)
        return a[0] if squeeze else a

    @property
    def launch_count(self):
        return sum(p.launch_count for p in (self.pre, self.gates, self.post) if p is not None)

    def reference_forward(self, state, h, c, emulate_bf16=True):
        """Plain PyTorch reference of one step: returns (action, h', c')."""
        x = torch.as_tensor(state, dtype=torch.float32, device=self.device)
        if self.pre:
            x = self.pre.reference_forward(x, emulate_bf16)
        g = self.gates.reference_forward(torch.cat([x, h], dim=1), emulate_bf16)
        H = self.H
        i, o, f, cc = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.sigmoid(g[:, 2 * H:3 * H]), torch.tanh(g[:, 3 * H:])
        c2 = f * c + i * cc
        h2 = o * torch.tanh(c2)
        a = self.post.reference_forward(h2, emulate_bf16) if self.post else h2.clamp(-1, 1)
        return a, h2, c2

    def close(self):
        for p in (self.pre, self.gates, self.post):
            if p is not None:
                p.close()


def build_policy(config, policy_path=None, state_dim=None, action_dim=None, device="cuda:0"):
    """core/policy.py:49-53.  MLP policies are read from the user's ONNX file (cosim_b200/onnx_reader.py); without a file the
    synthetic MLP of SURVEY.md section 8d is used.  LSTM policies (core/policy.py:24-47) are read from ONNX too."""
    if policy_path:
        from .onnx_reader import load_policy
        spec = load_policy(policy_path)              # the user's ONNX file, as in core/policy.py:7-9 / 26-29
        if bool(config.get("policy", {}).get("use_lstm")) != (spec["kind"] == "lstm"):
            raise RuntimeError("config['policy']['use_lstm'] does not match the ONNX graph")
        if spec["kind"] == "lstm":
            return LSTMPolicy(spec["lstm"], spec["pre"], spec["post"], spec["activation"], device, pre_activated=spec["pre_activated"])
        return MLPPolicy(spec["layers"], spec["activation"], device)
    if config.get("policy", {}).get("use_lstm"):
        raise RuntimeError("an LSTM policy needs an ONNX file")
    return MLPPolicy(synthetic_mlp(state_dim, action_dim), "elu", device)
