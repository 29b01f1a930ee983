"""Env-sharded vector environment: one process per GPU, `torch.distributed` plumbing.

B200-native replacement for the reference's `DistributedBreedGym`
(breedgym/vector/vec_env.py:150-236), which spawns one subprocess per device
and ships every observation back through host pipes.  Here each rank owns a
contiguous block of the E logical envs on its own GPU and steps it
independently; all constants (germplasm, thresholds, effects) are replicated
and the shared cross key / per-env reset keys are derived locally from the
same seed, so the union of the shards is bit-identical to one `VecBreedGym`
with E envs.  The only exchange is the all-gather of the per-env rewards
(float32[E/G] per rank).  On GPUs it needs no collective launch at all: the
kernel that reduces a step's GEBVs to rewards stores them straight into every
rank's receive window over NVLink peer memory (`PeerRewards`, csrc/peer.cu);
`bg_allgather_f32` (ncclAllGather through the C ABI) and `torch.distributed`
(gloo in the CPU tests) remain as alternatives.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .. import _lib
from .vec_env import VecBreedGym


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of `total` envs owned by `rank`: (begin, count); sizes differ by at most 1."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def shard_counts(total: int, world: int) -> List[int]:
    return [shard_range(total, world, r)[1] for r in range(world)]


def allgather_rewards(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Concatenate the per-rank reward vectors (`counts[r]` entries from rank r) on every rank."""
    world = len(counts)
    if world == 1:
        return local.clone()
    if len(set(counts)) == 1:
        out = torch.empty(sum(counts), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    cap = max(counts)
    padded = torch.zeros(cap, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty(cap, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


class RewardGather:
    """The path's one collective through the C ABI (`bg_allgather_f32`: ncclAllGather, no host round trip; the
    communicator is created and warmed once).  `torch.distributed` is used only to hand NCCL's unique id to the other
    ranks.  Equal shard sizes only (the sharded env pads otherwise through torch).

    The collective runs on its OWN stream: it waits (event) for the step that produced the local rewards, copies them
    into a private staging buffer and all-gathers from there, so the step stream never waits for the slowest rank --
    only whoever consumes the gathered rewards does.  `__call__(local, wait=True)` makes the step stream wait for the
    result right away (plain in-order semantics); `wait=False` leaves that to the caller (`self.done` is the event).
    """

    def __init__(self, simulator, world: int, rank: int, count: int, group=None):
        lib = _lib.load()
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(lib.bg_comm_unique_id(_lib.nptr(ident)))
        box = [ident.tobytes()]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = np.frombuffer(box[0], dtype=np.uint8).copy()
        self._comm = ctypes.c_void_p()
        _lib.check(lib.bg_comm_create(simulator._engine, _lib.nptr(ident), world, rank, ctypes.byref(self._comm)))
        self._fn = lib.bg_allgather_f32
        self._destroy = lib.bg_comm_destroy
        self.world, self.count = world, count
        self.device = simulator.device
        self.stream = torch.cuda.Stream(device=self.device)
        self._stream_ptr = ctypes.c_void_p(self.stream.cuda_stream)
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()
        # two staging / receive buffers, alternating: the rewards of an episode stay valid while the next episode runs
        self._stage = [torch.empty(count, dtype=torch.float32, device=self.device) for _ in range(2)]
        self._out = [torch.empty(world * count, dtype=torch.float32, device=self.device) for _ in range(2)]
        self._pos = 0

    def __call__(self, local: torch.Tensor, wait: bool = True) -> torch.Tensor:
        self._pos ^= 1
        stage, out = self._stage[self._pos], self._out[self._pos]
        main = torch.cuda.current_stream(self.device)
        self.ready.record(main)
        self.stream.wait_event(self.ready)
        with torch.cuda.stream(self.stream):
            stage.copy_(local, non_blocking=True)  # the step's reward buffer is free again as soon as this has run
        rc = self._fn(self._comm, stage.data_ptr(), out.data_ptr(), self.count, self._stream_ptr)
        if rc:
            _lib.check(rc)
        self.done.record(self.stream)
        if wait:
            main.wait_event(self.done)
        return out

    def close(self):
        if self._comm:
            self._destroy(self._comm)
            self._comm = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DeviceView:
    """`__cuda_array_interface__` over memory the library owns (a peer window)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


class PeerRewards:
    """The reward exchange over peer memory (`bg_peer_*`, csrc/peer.cu): once attached to the simulator's engine, every
    `bg_vec_step` that computes rewards also stores them into every rank's window -- no collective, no extra launch, no
    host work at an episode's end.  `handle` (bytes) is what the ranks exchange on the host; `connect(all_handles)`
    opens the windows (CUDA IPC between processes, raw pointers inside one process).

    `result()` is the window of the last published epoch as `float32[total]` (env order; this rank owns
    `[offset, offset + count)`, so ragged shards need no padding); it is complete once `wait()` (a
    one-warp kernel that spins on the arrival flags) has run on the consuming stream, and stays valid until this rank
    ends its NEXT episode.
    """

    def __init__(self, simulator, world: int, rank: int, total: int, offset: int):
        lib = _lib.load()
        self._lib = lib
        self._peer = ctypes.c_void_p()
        _lib.check(lib.bg_peer_create(simulator._engine, world, rank, total, offset, ctypes.byref(self._peer)))
        self._sim = simulator  # keeps the engine alive for as long as the exchange is attached to it
        self._engine = simulator._engine
        self.world, self.rank, self.total, self.offset = world, rank, total, offset
        self.device = simulator.device
        buf = np.zeros(_lib.PEER_HANDLE_BYTES, dtype=np.uint8)
        _lib.check(lib.bg_peer_handle(self._peer, _lib.nptr(buf)))
        self.handle = buf.tobytes()
        self._views = None
        self._wait_fn = lib.bg_peer_wait
        self._epoch_fn = lib.bg_peer_epoch
        self._raw_stream = torch._C._cuda_getCurrentRawStream
        self._dev_index = self.device.index

    def connect(self, handles: Sequence[bytes]):
        table = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
        if table.size != self.world * _lib.PEER_HANDLE_BYTES:
            raise ValueError("one handle per rank, in rank order")
        _lib.check(self._lib.bg_peer_connect(self._peer, _lib.nptr(table)))
        _lib.check(self._lib.bg_engine_set_peer(self._engine, self._peer))
        self._views = [torch.as_tensor(_DeviceView(int(self._lib.bg_peer_result(self._peer, par)), self.total), device=self.device)
                       for par in (0, 1)]

    def result(self) -> torch.Tensor:
        return self._views[self._epoch_fn(self._peer) & 1]

    def wait(self):
        """Order the current stream behind the arrival of every rank's rewards of the last epoch."""
        rc = self._wait_fn(self._peer, self._raw_stream(self._dev_index))
        if rc:
            _lib.check(rc)

    def publish(self, local: torch.Tensor):
        """Standalone publication of rewards that did not come out of `bg_vec_step` (host-mode steps)."""
        _lib.check(self._lib.bg_peer_publish_f32(self._peer, local.data_ptr(), local.numel(), self._raw_stream(self._dev_index)))

    def timeouts(self) -> int:
        return int(self._lib.bg_peer_timeouts(self._peer))

    def set_timeout_ms(self, ms: int):
        _lib.check(self._lib.bg_peer_set_timeout_ms(self._peer, int(ms)))

    def close(self):
        if self._peer:
            self._lib.bg_peer_destroy(self._peer)  # (detaches itself from the engine)
            self._peer = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedVecBreedGym:
    """`VecBreedGym` with `total_envs` logical envs partitioned over the ranks of a process group.

    `step(actions)` takes THIS rank's actions `[count, n, 2]` and returns the local
    observation handle with the rewards of ALL envs.  `collective`:
      "peer"   = the reward reduction stores into every rank's window over NVLink peer memory (`PeerRewards`,
                 csrc/peer.cu): no collective launch, no host work at an episode's end, ragged shards allowed;
      "native" = `bg_allgather_f32` (ncclAllGather through the C ABI, on its own stream; equal shards);
      "torch"  = `torch.distributed` (gloo in the CPU tests);
      "auto"   = "peer" when the group's backend is nccl and every rank could open every window, else "native"
                 (equal shards) or "torch".
    `async_rewards` (`info_device="device"`): the step stream does not wait for the other ranks' rewards -- the
    following steps overlap the exchange; call `env.wait_rewards()` on the stream that reads the returned rewards
    (they stay valid until this rank ends its next episode).

    `world` / `rank` override the process group's (several shards inside ONE process, e.g. two shards on one GPU in
    the tests): such shards are connected afterwards with `ShardedVecBreedGym.connect_local(shards)`.
    """

    def __init__(self, total_envs: int, group=None, device: Optional[int] = None, collective: str = "auto",
                 async_rewards: bool = False, world: Optional[int] = None, rank: Optional[int] = None, **kwargs):
        self.group = group
        local_only = world is not None
        if local_only:
            if rank is None or collective not in ("auto", "peer"):
                raise ValueError("world= needs rank= and the peer collective")
            self.world, self.rank = int(world), int(rank)
            collective = "peer"
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.total_envs = total_envs
        self.begin, self.count = shard_range(total_envs, self.world, self.rank)
        self.counts = shard_counts(total_envs, self.world)
        if device is None:
            device = torch.cuda.current_device()
        self.env = VecBreedGym(num_envs=self.count, env_shard=(self.begin, total_envs), device=device, **kwargs)
        self.num_envs = total_envs
        if collective not in ("auto", "peer", "native", "torch"):
            raise ValueError("collective must be 'auto', 'peer', 'native' or 'torch'")
        equal = len(set(self.counts)) == 1
        self._peer = None
        self._gather = None
        if self.world > 1 and collective in ("auto", "peer") and (local_only or dist.get_backend(group) == "nccl"):
            self._peer = self._make_peer(local_only, required=collective == "peer")
        if self._peer is not None:
            collective = "peer"
        elif collective == "peer" and self.world > 1:
            raise ValueError("the peer collective needs GPUs (nccl backend or in-process shards)")
        if collective == "auto":
            collective = "native" if (self.world > 1 and equal and dist.get_backend(group) == "nccl") else "torch"
        if collective == "native" and not equal:
            raise ValueError("the native reward all-gather needs equal shard sizes")
        self.collective = collective
        if collective == "native" and self.world > 1:
            self._gather = RewardGather(self.env.simulator, self.world, self.rank, self.count, group)
        self._zeros = None
        self._pin = None
        self.async_rewards = bool(async_rewards) and (self._gather is not None or self._peer is not None)

    # ---- peer windows ---------------------------------------------------------------
    def _make_peer(self, local_only: bool, required: bool):
        try:
            peer = PeerRewards(self.env.simulator, self.world, self.rank, self.total_envs, self.begin)
        except Exception:
            if required or local_only:
                raise
            peer = None
        if local_only:
            return peer  # connected later: ShardedVecBreedGym.connect_local
        # every rank learns every handle (host side, once) and whether every rank could open every window
        handles = [None] * self.world
        dist.all_gather_object(handles, peer.handle if peer is not None else None, group=self.group)
        ok = peer is not None and all(h is not None for h in handles)
        err = None
        if ok:
            try:
                peer.connect(handles)
            except Exception as e:  # e.g. no CUDA IPC between the ranks' containers
                ok, err = False, e
        oks = [None] * self.world
        dist.all_gather_object(oks, ok, group=self.group)
        if all(oks):
            return peer
        if peer is not None:
            peer.close()
        if required:
            raise RuntimeError(f"peer reward exchange unavailable on some rank ({err})")
        return None

    @staticmethod
    def connect_local(shards: Sequence["ShardedVecBreedGym"]):
        """Connect shards that live in ONE process (constructed with world= / rank=), in rank order."""
        shards = sorted(shards, key=lambda s: s.rank)
        handles = [s._peer.handle for s in shards]
        for s in shards:
            s._peer.connect(handles)

    @property
    def rewards_done(self):
        return self._gather.done if self._gather is not None else None

    def wait_rewards(self):
        """Make the current stream wait for the last reward exchange (no-op with the torch collective)."""
        if self._peer is not None:
            self._peer.wait()
        elif self._gather is not None:
            torch.cuda.current_stream(self.env.device).wait_event(self._gather.done)

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def local_slice(self) -> slice:
        return slice(self.begin, self.begin + self.count)

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        return self.env.reset(seed=seed, options=options)

    def gather_rewards(self, local: torch.Tensor) -> torch.Tensor:
        """float32[count] on this rank's GPU -> float32[total_envs] on every rank (enqueued, not synchronised)."""
        if self.world == 1:
            return local
        if self._peer is not None:  # rewards that did not come out of a step: publish them as an epoch of their own
            self._peer.publish(local.contiguous())
            if not self.async_rewards:
                self._peer.wait()
            return self._window()
        if self._gather is not None:
            return self._gather(local, wait=not self.async_rewards)
        return allgather_rewards(local, self.counts, self.group)

    def _window(self) -> torch.Tensor:
        """The last epoch's window, float32[total_envs] (a view of the library's memory, no copy)."""
        return self._peer.result()

    def step(self, local_actions):
        env = self.env
        will_reward = env.reward_shaping or env.step_idx + 1 == env.num_generations
        obs, rews, ter, tru, infos = env.step(local_actions)
        on_device = rews.__class__ is torch.Tensor  # info_device="device": rewards never leave the GPUs
        if will_reward:
            if self._peer is not None:
                # the step's own reduction has already stored this rank's rewards into every window
                if on_device:
                    if not self.async_rewards:
                        self._peer.wait()
                    rews = self._window()
                else:
                    self._peer.wait()
                    win = self._window()
                    if self._pin is None:
                        self._pin = torch.empty(self.total_envs, dtype=torch.float32, pin_memory=True)
                    self._pin.copy_(win, non_blocking=True)
                    torch.cuda.current_stream(env.device).synchronize()
                    rews = self._pin.numpy().copy()
            else:
                local = rews if on_device else torch.from_numpy(np.ascontiguousarray(rews, dtype=np.float32)).to(env.device)
                rews = self.gather_rewards(local)
                if not on_device:
                    rews = rews.cpu().numpy()
        elif on_device:
            if self._zeros is None:
                self._zeros = torch.zeros(self.total_envs, dtype=torch.float32, device=env.device)
            rews = self._zeros
        else:
            rews = np.zeros(self.total_envs)
        ter = np.full(self.total_envs, bool(ter[0]) if len(ter) else False)
        tru = np.full(self.total_envs, bool(tru[0]) if len(tru) else False)
        return obs, rews, ter, tru, infos

    def close(self):
        """Synchronise this rank's device and release the exchange (call it on every rank before the process group
        goes away; the receive window itself stays mapped until the process exits)."""
        torch.cuda.synchronize(self.env.device)
        if self._peer is not None:
            self._peer.close()
            self._peer = None
        if self._gather is not None:
            self._gather.close()
            self._gather = None
