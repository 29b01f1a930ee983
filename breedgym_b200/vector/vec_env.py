"""Vector environment (drop-in for `breedgym.vector.VecBreedGym`).

Mirrors breedgym/vector/vec_env.py:30-134: `E` independent populations stepped
together; ONE cross key per step shared by every env (the reference's
`jax.vmap(simulator.cross, in_axes=(None, 0))` runs the key split once, so all
envs see identical crossover masks -- reproduced on purpose); reward = max GEBV
over (individuals, traits) at the end of an episode (or every step with
`reward_shaping`); autoreset.

State lives on the GPU as bit planes `int32[E, n, 2, Wpad]`.  A step is one
C-ABI call (`bg_vec_step`): H2D of the actions, mask generation, blend, GEBV,
reward reduction, D2H of GEBV / rewards.

Multi-GPU (`env_shard=(begin, total)`): the E logical envs are partitioned into
contiguous blocks, one process per GPU; every shard derives the same cross key
and its own slice of the reset keys, so a sharded run is bit-identical to the
single-GPU env -- no data-path collective, only the reward all-gather
(`breedgym_b200.vector.sharded`).
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import Optional, Tuple, Union

import numpy as np
import torch

from .. import _lib
from ..gym_compat import VectorEnv, spaces
from ..population import PackedPopulation
from ..simulator import Simulator
from ..utils.paths import DATA_PATH

GENOME_FILE = DATA_PATH.joinpath("small_geno.npy")
_SIDE_STREAMS = {}  # device index -> the process-wide side stream of the reset prefetch


class VecBreedGym(VectorEnv):
    def __init__(
        self,
        num_envs: int = 1,
        initial_population: Union[str, Path, np.ndarray, PackedPopulation] = GENOME_FILE,
        individual_per_gen: Optional[int] = None,
        num_generations: int = 10,
        autoreset: bool = True,
        reward_shaping: bool = False,
        info_device: str = "host",
        env_shard: Optional[Tuple[int, int]] = None,
        **kwargs,
    ):
        self.num_envs = num_envs
        self.num_generations = num_generations
        self.autoreset = autoreset
        self.reward_shaping = reward_shaping
        if info_device not in ("host", "device"):
            raise ValueError("info_device must be 'host' or 'device'")
        self.info_device = info_device
        self.simulator = Simulator(**kwargs)
        self.device = self.simulator.device
        # logical env range of this shard: envs [begin, begin + num_envs) of `total`
        self.env_shard = (0, num_envs) if env_shard is None else (int(env_shard[0]), int(env_shard[1]))
        if self.env_shard[0] < 0 or self.env_shard[0] + num_envs > self.env_shard[1]:
            raise ValueError("env_shard=(begin, total) must contain [begin, begin + num_envs)")

        if isinstance(initial_population, (str, Path)):
            germplasm = self.simulator.load_population(initial_population)
        else:
            germplasm = self.simulator.as_packed(initial_population)
        self.germplasm = germplasm
        if individual_per_gen is None:
            individual_per_gen = len(self.germplasm)
        self.individual_per_gen = individual_per_gen
        self._set_spaces()

        self.populations = None
        self.step_idx = None
        self.reset_infos = {}
        self.random_key = None
        self._pinned = {}
        self._dev = {}
        self._io = None
        self._h2d_done = None
        self._vec_step_fn = _lib.load().bg_vec_step
        self._germ_gebv = None
        self._prefetched = None
        self._side = None

    def _set_spaces(self):
        n, m = self.individual_per_gen, self.germplasm.shape[1]
        self.single_observation_space = spaces.Box(low=0, high=1, shape=(n, m, 2), dtype=np.int8)
        self.single_action_space = spaces.Box(low=0, high=n, shape=(n, 2), dtype=np.int32)
        self.observation_space = spaces.Box(low=0, high=1, shape=(self.num_envs, n, m, 2), dtype=np.int8)
        self.action_space = spaces.Box(low=0, high=n, shape=(self.num_envs, n, 2), dtype=np.int32)

    # ---- buffers -------------------------------------------------------------------
    def _pinned_buf(self, name: str, shape, dtype) -> torch.Tensor:
        buf = self._pinned.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
            self._pinned[name] = buf
        return buf

    def _dev_buf(self, name: str, shape) -> torch.Tensor:
        buf = self._dev.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape):
            buf = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
            self._dev[name] = buf
        return buf

    def _zero_rewards(self) -> torch.Tensor:
        z = self._dev.get("zeros")
        if z is None or z.shape[0] != self.num_envs:
            z = self._dev["zeros"] = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        return z  # shared read-only tensor: intermediate steps carry no reward

    # ---- the hot path ----------------------------------------------------------------
    def cross(self, parents_idx) -> PackedPopulation:
        """Offspring of `populations[arange, parents_idx]`, one key for all envs (vec_env.py:75-77)."""
        return self.simulator.cross_envs(self.populations, parents_idx)

    def _host_io(self, shape, T):
        """Staging for the host-facing step, built once per action shape: pinned actions / GEBV / rewards
        (with their numpy views and raw pointers) and the device scratch behind them."""
        io = self._io
        if io is None or io["shape"] != shape:
            E, n = shape[0], shape[1]
            dev = self.device
            act_pin = torch.empty(shape, dtype=torch.int32, pin_memory=True)
            gebv_pin = torch.empty((E, n, T), dtype=torch.float32, pin_memory=True)
            rew_pin = torch.empty((E,), dtype=torch.float32, pin_memory=True)
            act_dev = torch.empty(shape, dtype=torch.int32, device=dev)
            gebv_dev = torch.empty((E, n, T), dtype=torch.float32, device=dev)
            rew_dev = torch.empty((E,), dtype=torch.float32, device=dev)
            io = self._io = {
                "shape": shape, "keep": (act_pin, gebv_pin, rew_pin, act_dev, gebv_dev, rew_dev),
                "act_np": act_pin.numpy(), "gebv_np": gebv_pin.numpy(), "rew_np": rew_pin.numpy(),
                "act_pin": act_pin.data_ptr(), "gebv_pin": gebv_pin.data_ptr(), "rew_pin": rew_pin.data_ptr(),
                "act_dev": act_dev.data_ptr(), "gebv_dev": gebv_dev.data_ptr(), "rew_dev": rew_dev.data_ptr(),
            }
        return io

    def step(self, actions):
        sim, E, T = self.simulator, self.num_envs, self.simulator.GEBV_model.n_traits
        done = self.step_idx + 1 == self.num_generations
        need_reward = self.reward_shaping or done
        host_info = self.info_device == "host"
        src = self.populations.words
        n_src = src.shape[1]
        on_device = isinstance(actions, torch.Tensor) and actions.is_cuda

        if host_info and not on_device:
            # ---- the Gym-facing path: numpy actions in, numpy GEBV / rewards out, one sync inside the C call
            a = np.asarray(actions)
            if a.ndim != 3 or a.shape[0] != E or a.shape[2] != 2:
                raise ValueError(f"actions must have shape ({E}, n, 2), got {a.shape}")
            n = a.shape[1]
            io = self._io
            if io is None or io["shape"] != a.shape:
                io = self._host_io(a.shape, T)
            io["act_np"][...] = a  # int64 -> int32 conversion happens in this copy
            out = torch.empty((E, n, 2, sim.words_per_row), dtype=torch.int32, device=self.device)
            # random_key, k = split(random_key): advances the chain; k and the next k sit in sim._chain_out
            rc = sim._chain_fn(sim._key_ptr, sim._layout_id, sim._chain_ptr)
            if rc:
                _lib.check(rc)
            kp = sim._chain_out_addr
            rc = self._vec_step_fn(
                sim._engine, src.data_ptr(), out.data_ptr(), io["act_pin"], io["act_dev"], E, n_src, n, kp, kp + 8,
                sim._layout_id, sim._schedule_id, io["gebv_dev"], io["rew_dev"] if need_reward else None,
                io["gebv_pin"], io["rew_pin"] if need_reward else None, sim._stream())
            if rc:
                _lib.check(rc)
            infos = {"GEBV": io["gebv_np"].copy()}
            rews = io["rew_np"].copy() if need_reward else np.zeros(E)
        else:
            if on_device:
                act_dev = actions.to(device=self.device, dtype=torch.int32).contiguous()
                act_host_ptr = None
            else:
                a = np.asarray(actions)
                act_pin = self._pinned_buf("actions", a.shape, torch.int32)
                if self._h2d_done is not None:  # previous async H2D must have left the staging buffer
                    self._h2d_done.synchronize()
                act_pin.numpy()[...] = a
                act_dev = torch.empty(a.shape, dtype=torch.int32, device=self.device)
                act_host_ptr = act_pin.data_ptr()
            if act_dev.dim() != 3 or act_dev.shape[0] != E or act_dev.shape[2] != 2:
                raise ValueError(f"actions must have shape ({E}, n, 2), got {tuple(act_dev.shape)}")
            n = act_dev.shape[1]
            out = sim._empty_words(E, n)
            if host_info:
                gebv_dev = self._dev_buf("gebv", (E, n, T))
                rew_dev = self._dev_buf("rews", (E,)) if need_reward else None
                gebv_pin = self._pinned_buf("gebv", (E, n, T), torch.float32)
                rew_pin = self._pinned_buf("rews", (E,), torch.float32) if need_reward else None
            else:
                gebv_dev = torch.empty((E, n, T), dtype=torch.float32, device=self.device)
                rew_dev = torch.empty((E,), dtype=torch.float32, device=self.device) if need_reward else None
                gebv_pin = rew_pin = None
            sim._next_key(lookahead=True)
            kp = sim._chain_out.ctypes.data
            _lib.check(self._vec_step_fn(
                sim._engine, src.data_ptr(), out.data_ptr(), act_host_ptr, act_dev.data_ptr(), E, n_src, n, kp, kp + 8,
                sim._layout_id, sim._schedule_id, gebv_dev.data_ptr(), rew_dev.data_ptr() if need_reward else None,
                gebv_pin.data_ptr() if host_info else None, rew_pin.data_ptr() if rew_pin is not None else None,
                sim._stream()))
            if act_host_ptr is not None and not host_info:  # no sync happened inside the call
                self._h2d_done = torch.cuda.Event()
                self._h2d_done.record(torch.cuda.current_stream(self.device))
            else:
                self._h2d_done = None
            if host_info:
                infos = {"GEBV": gebv_pin.numpy().copy()}
                rews = rew_pin.numpy().copy() if need_reward else np.zeros(E)
            else:  # device mode: nothing leaves the GPU, nothing synchronises
                infos = {"GEBV": gebv_dev}
                rews = rew_dev if need_reward else self._zero_rewards()

        self.populations = PackedPopulation._trusted(sim, out)
        self.step_idx += 1
        if done and self.autoreset:
            self.reset()
        return self.populations, rews, np.zeros(E, dtype=bool), np.full(E, done), infos

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        self.step_idx = 0
        if seed is not None:
            self.simulator.set_seed(seed)
            self.random_key = _lib.key_data(seed)
        elif self.random_key is None:
            seed = np.random.randint(2**32)
            self.random_key = _lib.key_data(seed)

        if options is not None and "individual_per_gen" in options.keys():
            self.individual_per_gen = options["individual_per_gen"]
            self._set_spaces()

        begin, total = self.env_shard
        sim, E, n = self.simulator, self.num_envs, self.individual_per_gen
        T = sim.GEBV_model.n_traits
        key = np.ascontiguousarray(self.random_key, dtype=np.uint32)
        if self._germ_gebv is None and not os.environ.get("BG_NO_GERM_GEBV"):  # once: the reset infos are gathered from the germplasm's GEBVs
            self._germ_gebv = sim._gebv(self.germplasm).to(torch.float32).contiguous()  # raw kernel output, as bg_vec_reset computes
        host_info = self.info_device == "host"
        pre, self._prefetched = self._prefetched, None
        if pre is not None and np.array_equal(pre["key"], key) and pre["n"] == n:
            # this reset was drawn ahead of time on the side stream (see _prefetch_reset): adopt its buffers
            main = torch.cuda.current_stream(self.device)
            main.wait_event(pre["event"])
            idx, words, gebv_dev = pre["idx"], pre["words"], pre["gebv_dev"]
            if host_info:
                pre["event"].synchronize()
                infos_gebv = pre["gebv_pin"].numpy().copy()
            else:
                infos_gebv = gebv_dev
        else:
            idx = torch.empty((E, n), dtype=torch.int32, device=self.device)
            words = sim._empty_words(E, n)
            germ = self.germplasm.words.contiguous()
            if host_info:
                io = self._host_io((E, n, 2), T)
                gebv_dev_ptr, gebv_host_ptr = io["gebv_dev"], io["gebv_pin"]
            else:
                gebv_dev = torch.empty((E, n, T), dtype=torch.float32, device=self.device)
                gebv_dev_ptr, gebv_host_ptr = gebv_dev.data_ptr(), None
            # env g draws permutation(keys[1 + g], N)[:n] with keys = split(random_key, total + 1); one C call does
            # the draw, the gather from the germplasm and the reset infos
            _lib.check(_lib.load().bg_vec_reset(sim._engine, germ.data_ptr(), germ.shape[0], _lib.nptr(key), total, begin, E, n,
                                                sim._layout_id, idx.data_ptr(), words.data_ptr(), gebv_dev_ptr, gebv_host_ptr,
                                                self._germ_gebv.data_ptr() if self._germ_gebv is not None else None, sim._stream()))
            infos_gebv = io["gebv_np"].copy() if host_info else gebv_dev
        self.random_key = _lib.key_split_at(self.random_key, 0, total + 1, sim.rng_layout)
        self._reset_indices = idx
        self.populations = PackedPopulation(sim, words)
        self.reset_infos = {"GEBV": infos_gebv}
        # Opt-in experiment (BG_RESET_PREFETCH=1, device mode): draw the next reset ahead of time on a side stream.
        # +5 % env-steps/s when it works, but whole runs at half speed now and then (4 of 6; cause not found: neither
        # a shared side stream nor more hardware connections cure it), so it is off by default; with host infos its stream / event
        # bookkeeping costs more host time than the reset kernels it hides (640 k -> 580 k env-steps/s end to end).
        if (self.autoreset and self.info_device == "device" and self._germ_gebv is not None
                and os.environ.get("BG_RESET_PREFETCH")):
            self._prefetch_reset()
        return self.populations, self.reset_infos

    def _prefetch_reset(self):
        """Draw the NEXT reset now, on a side stream, while the episode runs: it depends on `random_key` and the
        germplasm only (vec_env.py:109-130), so at the end of the episode the autoreset adopts finished buffers instead
        of running the permutation + gather (+ copy of the infos) on the step's critical path.  A reset with a seed, other
        options or a `random_key` somebody changed in between simply ignores it."""
        sim, E, n = self.simulator, self.num_envs, self.individual_per_gen
        begin, total = self.env_shard
        T = sim.GEBV_model.n_traits
        if self._side is None:
            # ONE side stream per device, shared by all envs of the process: every extra stream takes one of the (8 by
            # default) hardware connections, and streams that have to share one serialise against each other
            key_dev = (self.device.index if self.device.index is not None else torch.cuda.current_device())
            if key_dev not in _SIDE_STREAMS:
                _SIDE_STREAMS[key_dev] = torch.cuda.Stream(device=self.device)
            self._side = _SIDE_STREAMS[key_dev]
        key = np.array(self.random_key, dtype=np.uint32, copy=True)
        germ = self.germplasm.words.contiguous()
        host_info = self.info_device == "host"
        main = torch.cuda.current_stream(self.device)
        # buffers from the MAIN stream's pool (the pool every step allocates its 60 MB population from), handed to the
        # side stream with record_stream: allocating them under the side stream made the caching allocator fall back
        # to synchronising cudaMalloc / cudaFree now and then (whole runs at 0.75 M instead of 1.36 M env-steps/s)
        idx = torch.empty((E, n), dtype=torch.int32, device=self.device)
        words = sim._empty_words(E, n)
        gebv_dev = torch.empty((E, n, T), dtype=torch.float32, device=self.device)
        for t in (idx, words, gebv_dev):
            t.record_stream(self._side)
        self._side.wait_stream(main)  # the buffers' previous users, and the germplasm GEBVs computed on the main stream
        _lib.check(_lib.load().bg_vec_reset(sim._engine, germ.data_ptr(), germ.shape[0], _lib.nptr(key), total, begin, E, n,
                                            sim._layout_id, idx.data_ptr(), words.data_ptr(), gebv_dev.data_ptr(), None,
                                            self._germ_gebv.data_ptr(), ctypes.c_void_p(self._side.cuda_stream)))
        gebv_pin = None
        if host_info:
            gebv_pin = self._pinned_buf("reset_gebv", (E, n, T), torch.float32)
            with torch.cuda.stream(self._side):
                gebv_pin.copy_(gebv_dev, non_blocking=True)
        event = torch.cuda.Event()
        event.record(self._side)
        self._prefetched = {"key": key, "n": n, "idx": idx, "words": words, "gebv_dev": gebv_dev, "gebv_pin": gebv_pin,
                            "event": event}

    def get_info(self) -> dict:
        gebv = self.simulator.GEBV_model(self.populations)
        return {"GEBV": gebv.cpu().numpy() if self.info_device == "host" else gebv}

    def set_attr(self, name, values):
        return setattr(self, name, values)


class _VecBreedGym(VecBreedGym):
    """Shard worker of the reference's DistributedBreedGym (vec_env.py:140-147): scalar ter/tru."""

    def step(self, action):
        obs, rews, ter, tru, infos = super().step(action)
        assert np.all(ter == ter[0])
        assert np.all(tru == tru[0])
        return obs, rews, ter[0], tru[0], infos
