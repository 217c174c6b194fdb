"""On-device policy plumbing for BASELINE config 5 (self-play rollout with a TD3 actor) -- SURVEY.md section 8f rank 2.

The actor is the reference's `ActorNetwork` (rl/td3/networks.py:6-20: 18 -> 256 -> 256 -> 4, tanh after every layer);
it consumes the env's observation tensor in place (no host round trip) and its output feeds `HockeyVecEnv.step`.
The dense layers are plain library matmuls (torch/cuBLAS): they are the op adjacent to the hot path, not the path.
"""
import torch
from torch import nn


class ActorNetwork(nn.Module):
    def __init__(self, obs_dim=18, action_dim=4, hidden=256):
        super().__init__()
        self.fc1 = nn.Linear(obs_dim, hidden)
        self.fc2 = nn.Linear(hidden, hidden)
        self.fc3 = nn.Linear(hidden, action_dim)

    def forward(self, obs):
        x = torch.tanh(self.fc1(obs))
        x = torch.tanh(self.fc2(x))
        return torch.tanh(self.fc3(x))


def load_td3_actor(path, device="cuda:0", name=None):
    """The reference's trained policy as an on-device module.  `path` is either a reference checkpoint written by
    `TD3Agent.save` (rl/td3/agent.py:269-275: a dict whose `policy` entry is the ActorNetwork state_dict) or an
    `.npz` of float32 arrays `<name>.fc1.weight` ... `<name>.fc3.bias` (tests/golden/td3_actors.npz, extracted from
    those checkpoints by tests/golden/make_actor_fixtures.py)."""
    if str(path).endswith(".npz"):
        import numpy as np
        z = np.load(path)
        names = sorted({k.split(".")[0] for k in z.files})
        if name is None:
            if len(names) != 1:
                raise ValueError(f"{path} holds several actors {names}: pass name=")
            name = names[0]
        if name not in names:
            raise ValueError(f"no actor {name!r} in {path}; have {names}")
        sd = {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + ".")}
    else:
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        sd = ckpt["policy"] if "policy" in ckpt else ckpt
    hidden, obs_dim = sd["fc1.weight"].shape
    actor = ActorNetwork(obs_dim=obs_dim, action_dim=sd["fc3.weight"].shape[0], hidden=hidden)
    actor.load_state_dict(sd)
    return actor.to(device).eval()


class FusedActor:
    """The same network as ONE fused sm_100a tensor-core kernel (csrc/hk_actor.cuh, hk_actor_forward): weights resident in
    shared memory (layer 1 TF32, layers 2-3 bf16), fp32 accumulation in tensor memory, hidden activations never leave the SM.  Callable like the
    module: obs [N,18] float32 CUDA tensor -> actions [N,4] (a persistent output buffer, overwritten by the next call;
    pass `out=` to write into columns 0..3 of an existing [N,4] / [N,8] action tensor).  There is no fallback: without the
    CUDA library or a GPU this raises."""

    def __init__(self, actor, device="cuda:0"):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib
        self.L = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.HockeyLibraryError("FusedActor needs a CUDA device: the actor kernel has no CPU fallback")
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in actor.state_dict().items()}
        if tuple(sd["fc1.weight"].shape) != (256, 18) or tuple(sd["fc2.weight"].shape) != (256, 256) or tuple(sd["fc3.weight"].shape) != (4, 256):
            raise ValueError("FusedActor implements the reference architecture 18 -> 256 -> 256 -> 4")
        self.params = self.pack(sd).to(self.device)
        assert self.params.numel() == int(self.L.hk_actor_param_bytes())
        self._out = None

    @staticmethod
    def _core_matrix_order(w, n_pad, k_pad, dtype):
        """[N, K] float weight -> bytes in UMMA K-major core-matrix order (8-row x 16-byte core matrices, no swizzle):
        with e = 16 / itemsize elements per 16-byte chunk, element (n, k) sits at
        (k // e) * (n_pad * 16) + n * 16 + (k % e) * itemsize bytes.  dtype bfloat16, or float32 for the TF32 layer."""
        full = torch.zeros((n_pad, k_pad), dtype=torch.float32)
        full[:w.shape[0], :w.shape[1]] = w
        e = 16 // torch.empty((), dtype=dtype).element_size()
        return full.to(dtype).view(n_pad, k_pad // e, e).permute(1, 0, 2).contiguous().view(torch.uint8).flatten()

    @classmethod
    def pack(cls, sd):
        b3 = torch.zeros(16, dtype=torch.float32)
        b3[:4] = sd["fc3.bias"]
        parts = [cls._core_matrix_order(sd["fc1.weight"], 256, 24, torch.float32),
                 cls._core_matrix_order(sd["fc2.weight"], 256, 256, torch.bfloat16),
                 cls._core_matrix_order(sd["fc3.weight"], 16, 256, torch.bfloat16),
                 sd["fc1.bias"].contiguous().view(torch.uint8).flatten(),
                 sd["fc2.bias"].contiguous().view(torch.uint8).flatten(), b3.view(torch.uint8).flatten()]
        return torch.cat(parts).contiguous()

    def __call__(self, obs, out=None):
        if not (obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and obs.dim() == 2 and obs.shape[1] == 18):
            raise ValueError("FusedActor expects a contiguous float32 CUDA tensor [N, 18]")
        n = obs.shape[0]
        if out is None:
            if self._out is None or self._out.shape[0] != n:
                self._out = torch.empty((n, 4), dtype=torch.float32, device=obs.device)
            out = self._out
        if not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 2 and out.shape[0] == n and out.stride(1) == 1 and out.shape[1] >= 4):
            raise ValueError("out must be a float32 CUDA tensor [N, >= 4] with unit column stride")
        C = self._C
        with torch.cuda.device(obs.device):
            self._lib.check(self.L.hk_actor_forward(C.c_void_p(self.params.data_ptr()), C.c_void_p(obs.data_ptr()),
                                                    C.c_void_p(out.data_ptr()), int(out.stride(0)), n, obs.device.index or 0,
                                                    C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)))
        return out[:, :4] if out.shape[1] != 4 else out


@torch.no_grad()
def actor_rollout(env, actor, steps, opponent_actor=None):
    """`steps` ticks of `env` (a HockeyVecEnv with p1 external) driven by `actor`; player 2 is the env's in-kernel
    BasicOpponent, or `opponent_actor` acting on obs_agent_two() (the PolicyOpponent pattern, hockey_env.py:908-922)
    when the env was built with p2=None.  Returns the env's episode statistics."""
    obs = env.obs
    for _ in range(steps):
        a1 = actor(obs)
        if opponent_actor is not None:
            a2 = opponent_actor(env.obs_agent_two())
            obs, *_ = env.step(torch.cat([a1, a2], dim=1).contiguous())
        else:
            obs, *_ = env.step(a1.contiguous())
    return env.stats()
