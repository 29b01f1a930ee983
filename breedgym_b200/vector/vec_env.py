"""Vector environment (drop-in for `breedgym.vector.VecBreedGym`).

Mirrors breedgym/vector/vec_env.py:30-134: `E` independent populations stepped
together; ONE cross key per step shared by every env (the reference's
`jax.vmap(simulator.cross, in_axes=(None, 0))` runs the key split once, so all
envs see identical crossover masks -- reproduced on purpose); reward = max GEBV
over (individuals, traits) at the end of an episode (or every step with
`reward_shaping`); autoreset.

State lives on the GPU as bit planes `int32[E, n, 2, Wpad]`.  A step is ONE
C-ABI call (`bg_vec_step`): H2D of the actions, key-chain advance, the fused
cross + GEBV kernel, reward reduction, D2H of GEBV / rewards; the crossover
masks of the following steps are generated ahead of time on a side stream.  The
Python side of a step allocates nothing in `info_device="device"` mode: the
observations cycle through `obs_ring` preallocated buffers (so an observation
handle stays valid for `obs_ring - 2` further steps; copy it to keep it longer).

Multi-GPU (`env_shard=(begin, total)`): the E logical envs are partitioned into
contiguous blocks, one process per GPU; every shard derives the same cross key
and its own slice of the reset keys, so a sharded run is bit-identical to the
single-GPU env -- no data-path collective, only the reward all-gather
(`breedgym_b200.vector.sharded`).
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple, Union

import numpy as np
import torch

from .. import _lib
from ..gym_compat import VectorEnv, spaces
from ..population import PackedPopulation, ParentsView
from ..simulator import Simulator
from ..utils.paths import DATA_PATH

GENOME_FILE = DATA_PATH.joinpath("small_geno.npy")


class _ObsRing:
    """`slots` preallocated (population, GEBV, reward) buffer sets for one (E, n): the device-mode step writes into the
    next one instead of allocating (a 60 MB `torch.empty` per step costs host time, and a cudaMalloc when the caching
    allocator has no free block -- e.g. at the first autoreset -- stalls the stream for milliseconds)."""

    def __init__(self, sim: Simulator, E: int, n: int, T: int, slots: int):
        dev = sim.device
        self.n = n
        self.slots = slots
        self.pos = 0
        self.words = [sim._empty_words(E, n) for _ in range(slots)]
        self.pop_ptr = [w.data_ptr() for w in self.words]
        self.gebv = [torch.empty((E, n, T), dtype=torch.float32, device=dev) for _ in range(slots)]
        self.gebv_ptr = [g.data_ptr() for g in self.gebv]
        self.rew = [torch.empty((E,), dtype=torch.float32, device=dev) for _ in range(slots)]
        self.rew_ptr = [r.data_ptr() for r in self.rew]
        self.infos = [{"GEBV": g} for g in self.gebv]

    def take(self, src_ptr: int) -> int:
        """Next slot that does not hold the population about to be read."""
        pos = self.pos + 1
        if pos == self.slots:
            pos = 0
        if self.pop_ptr[pos] == src_ptr:
            pos = pos + 1 if pos + 1 < self.slots else 0
        self.pos = pos
        return pos


class _LazyInfos(dict):
    """`reset_infos` of an AUTORESET in host mode: the GEBVs of the new populations are computed on the GPU in stream
    order, but copied to the host only if somebody reads them (the step that triggered the autoreset returns the infos of
    the final offspring, not these: vec_env.py:102-107) -- no extra copy + synchronisation inside that step."""

    def __init__(self, fetch):
        super().__init__()
        self._fetch = fetch

    def _materialise(self):
        if self._fetch is not None:
            fetch, self._fetch = self._fetch, None
            super().__setitem__("GEBV", fetch())

    def __getitem__(self, key):
        self._materialise()
        return super().__getitem__(key)

    def __contains__(self, key):
        return key == "GEBV" or super().__contains__(key)

    def keys(self):
        self._materialise()
        return super().keys()

    def items(self):
        self._materialise()
        return super().items()

    def values(self):
        self._materialise()
        return super().values()

    def get(self, key, default=None):
        self._materialise()
        return super().get(key, default)

    def __iter__(self):
        self._materialise()
        return super().__iter__()

    def __len__(self):
        return 1 if self._fetch is not None else super().__len__()


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
        obs_ring: Optional[int] = None,
        reset_prefetch: Optional[bool] = None,
        **kwargs,
    ):
        self.num_envs = num_envs
        self.num_generations = num_generations
        self.autoreset = autoreset
        self.reward_shaping = reward_shaping
        if info_device not in ("host", "device"):
            raise ValueError("info_device must be 'host' or 'device'")
        self.info_device = info_device
        # observation buffers: "device" mode cycles through 3 preallocated ones, "host" mode (the Gym-facing default)
        # hands out a fresh buffer per step like the reference does; obs_ring overrides (0 = fresh, >= 3 = ring)
        if obs_ring is None:
            obs_ring = 3 if info_device == "device" else 0
        if obs_ring != 0 and obs_ring < 3:
            raise ValueError("obs_ring must be 0 (a fresh observation buffer per step) or >= 3")
        self.obs_ring = obs_ring
        # the NEXT autoreset is drawn on the library's side stream while the episode runs (it depends on the reset key
        # chain and the germplasm only) and adopted at the episode's end
        self.reset_prefetch = True if reset_prefetch is None else bool(reset_prefetch)
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
        self._io = None
        self._ring = None
        self._own = None           # the population object this env produced last (trusted layout)
        self._h2d_done = None
        self._idx_buf = None
        self._germ_words = None
        self._germ_gebv = None
        self.reuse_germplasm_gebv = True  # reset infos gathered from the germplasm's GEBVs (bit-identical to re-scoring)
        self._pre = None           # prefetched reset: two buffer sets, alternating
        lib = _lib.load()
        self._vec_step_fn = lib.bg_vec_step
        self._vec_reset_fn = lib.bg_vec_reset
        self._prefetch_fn = lib.bg_vec_reset_prefetch
        self._adopt_fn = lib.bg_vec_reset_adopt
        self._raw_stream = torch._C._cuda_getCurrentRawStream
        self._dev_index = self.device.index

    def _set_spaces(self):
        n, m = self.individual_per_gen, self.germplasm.shape[1]
        self.single_observation_space = spaces.Box(low=0, high=1, shape=(n, m, 2), dtype=np.int8)
        self.single_action_space = spaces.Box(low=0, high=n, shape=(n, 2), dtype=np.int32)
        self.observation_space = spaces.Box(low=0, high=1, shape=(self.num_envs, n, m, 2), dtype=np.int8)
        self.action_space = spaces.Box(low=0, high=n, shape=(self.num_envs, n, 2), dtype=np.int32)

    # ---- buffers -------------------------------------------------------------------
    def _host_io(self, shape, T):
        """Staging for host-side actions / infos, built once per action shape: pinned actions / GEBV / rewards (with
        their numpy views and raw pointers) and the device scratch behind them."""
        io = self._io
        if io is None or io["shape"] != shape:
            if self._h2d_done is not None:
                self._h2d_done.synchronize()
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

    def _obs_slots(self, n: int, T: int) -> Optional[_ObsRing]:
        if not self.obs_ring:
            return None
        ring = self._ring
        if ring is None or ring.n != n:
            ring = self._ring = _ObsRing(self.simulator, self.num_envs, n, T, self.obs_ring)
        return ring

    def _source_words(self) -> torch.Tensor:
        """Bit planes of `self.populations`, validated unless this env produced them (wrappers and users may assign
        arrays, tensors or populations of other sizes to `env.populations`, as the reference allows)."""
        pop = self.populations
        if pop is self._own:
            return pop.words
        pop = self.simulator.as_packed(pop)
        words = pop.words
        if words.dim() != 4 or words.shape[0] != self.num_envs:
            raise ValueError(f"populations must have shape ({self.num_envs}, n, m, 2), got {tuple(pop.shape)}")
        if not words.is_contiguous():
            pop = PackedPopulation(self.simulator, words.contiguous())
        self.populations = self._own = pop
        return pop.words

    # ---- the hot path ----------------------------------------------------------------
    def cross(self, parents) -> PackedPopulation:
        """Offspring of `populations[arange, parents_idx]`, one key for all envs (vec_env.py:75-77).  Takes the index
        pairs `int[E, n, 2]`, or the lazy view the reference's idiom `populations[arange(E)[:, None, None], actions]`
        returns here."""
        if isinstance(parents, ParentsView):
            return self.simulator.cross(parents)
        return self.simulator.cross_envs(self.populations, parents)

    def step(self, actions):
        sim, E = self.simulator, self.num_envs
        T = sim.GEBV_model.n_traits
        done = self.step_idx + 1 == self.num_generations
        need_reward = self.reward_shaping or done
        host_info = self.info_device == "host"
        src = self._source_words()
        src_ptr = src.data_ptr()

        # ---- actions: a CUDA tensor is used in place, anything else goes through the pinned staging buffer
        if actions.__class__ is torch.Tensor and actions.is_cuda:
            if actions.dtype != torch.int32 or not actions.is_contiguous() or actions.device != self.device:
                actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
            shape = tuple(actions.shape)
            act_host_ptr, act_dev_ptr = None, actions.data_ptr()
            io = None
        elif (actions.__class__ is torch.Tensor and actions.dtype == torch.int32 and actions.is_contiguous()
              and actions.is_pinned()):
            # a pinned host tensor goes to the device as it is: no staging copy (the asynchronous H2D reads it, so in
            # device mode the caller keeps it unchanged until the step has run; host mode synchronises anyway)
            shape = tuple(actions.shape)
            io = self._io
            if io is None or io["shape"] != shape:
                if len(shape) != 3 or shape[0] != E or shape[2] != 2:
                    raise ValueError(f"actions must have shape ({E}, n, 2), got {shape}")
                io = self._host_io(shape, T)
            act_host_ptr, act_dev_ptr = actions.data_ptr(), io["act_dev"]
        else:
            a = np.asarray(actions)
            shape = a.shape
            io = self._io
            if io is None or io["shape"] != shape:
                if a.ndim != 3 or shape[0] != E or shape[2] != 2:
                    raise ValueError(f"actions must have shape ({E}, n, 2), got {shape}")
                io = self._host_io(shape, T)
            if self._h2d_done is not None:  # an earlier asynchronous H2D must have left the staging buffer
                self._h2d_done.synchronize()
            io["act_np"][...] = a  # int64 -> int32 conversion happens in this copy
            act_host_ptr, act_dev_ptr = io["act_pin"], io["act_dev"]
        if len(shape) != 3 or shape[0] != E or shape[2] != 2:
            raise ValueError(f"actions must have shape ({E}, n, 2), got {shape}")
        n = shape[1]

        # ---- outputs: a ring slot, or fresh buffers
        ring = self._ring
        if self.obs_ring and (ring is None or ring.n != n):
            ring = self._obs_slots(n, T)
        if ring is not None:
            slot = ring.take(src_ptr)
            # (a fresh handle object per step: a bool view somebody materialised from an older handle stays a valid snapshot)
            out_pop, out_ptr = PackedPopulation._trusted(sim, ring.words[slot]), ring.pop_ptr[slot]
        else:
            out_words = torch.empty((E, n, 2, sim.words_per_row), dtype=torch.int32, device=self.device)
            out_pop, out_ptr = PackedPopulation._trusted(sim, out_words), out_words.data_ptr()
        if host_info:
            if io is None or io["shape"] != shape:
                io = self._host_io(shape, T)
            gebv_ptr, rew_ptr = io["gebv_dev"], io["rew_dev"] if need_reward else None
            gebv_host, rew_host = io["gebv_pin"], io["rew_pin"] if need_reward else None
        elif ring is not None:
            gebv_ptr, rew_ptr = ring.gebv_ptr[slot], ring.rew_ptr[slot] if need_reward else None
            gebv_host = rew_host = None
        else:
            gebv_t = torch.empty((E, n, T), dtype=torch.float32, device=self.device)
            rew_t = torch.empty((E,), dtype=torch.float32, device=self.device) if need_reward else None
            gebv_ptr, rew_ptr = gebv_t.data_ptr(), rew_t.data_ptr() if need_reward else None
            gebv_host = rew_host = None

        # ---- ONE library call: H2D, `random_key, k = split(random_key)`, cross + GEBV, reward, D2H (+ sync iff D2H)
        rc = self._vec_step_fn(sim._engine, src_ptr, out_ptr, act_host_ptr, act_dev_ptr, E, src.shape[1], n, sim._key_ptr,
                               sim._layout_id, sim._schedule_id, gebv_ptr, rew_ptr, gebv_host, rew_host,
                               self._raw_stream(self._dev_index))
        if rc:
            _lib.check(rc)

        if host_info:
            infos = {"GEBV": io["gebv_np"].copy()}
            rews = io["rew_np"].copy() if need_reward else np.zeros(E)
        else:  # device mode: nothing leaves the GPU, nothing synchronises
            if act_host_ptr is not None:
                if self._h2d_done is None:
                    self._h2d_done = torch.cuda.Event()
                self._h2d_done.record(torch.cuda.current_stream(self.device))
            if ring is not None:
                infos = ring.infos[slot]
                rews = ring.rew[slot] if need_reward else self._zero_rewards()
            else:
                infos = {"GEBV": gebv_t}
                rews = rew_t if need_reward else self._zero_rewards()

        self.populations = self._own = out_pop
        self.step_idx += 1
        if done and self.autoreset:
            self.reset(_auto=True)
        return self.populations, rews, np.zeros(E, dtype=bool), np.full(E, done), infos

    def _zero_rewards(self) -> torch.Tensor:
        z = getattr(self, "_zeros", None)
        if z is None or z.shape[0] != self.num_envs:
            z = self._zeros = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        return z  # shared read-only tensor: intermediate steps carry no reward

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None, _auto: bool = False):
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
        if self._germ_words is None:
            self._germ_words = self.germplasm.words.contiguous()
        germ = self._germ_words
        if self._germ_gebv is None and self.reuse_germplasm_gebv:
            # once: the reset infos are gathered from the germplasm's GEBVs (raw kernel output, as bg_vec_reset computes)
            self._germ_gebv = sim._gebv(self.germplasm).to(torch.float32).contiguous()
        germ_gebv_ptr = self._germ_gebv.data_ptr() if (self._germ_gebv is not None and self.reuse_germplasm_gebv) else None
        host_info = self.info_device == "host"
        idx = self._idx_buf
        if idx is None or tuple(idx.shape) != (E, n):
            idx = self._idx_buf = torch.empty((E, n), dtype=torch.int32, device=self.device)

        pre = self._pre
        if (pre is not None and pre["pending"] and pre["n"] == n and pre["E"] == E
                and pre["key"][0] == key[0] and pre["key"][1] == key[1]):
            # this reset was drawn ahead of time on the side stream: the step stream waits for it and adopts its buffers
            rc = self._adopt_fn(sim._engine, self._raw_stream(self._dev_index))
            if rc:
                _lib.check(rc)
            pre["pending"] = False
            words, gebv_t, idx = pre["bufs"][pre["slot"]]
            self.random_key = _lib.key_split_at(self.random_key, 0, total + 1, sim.rng_layout)
            self._reset_indices = idx
            self.populations = self._own = PackedPopulation._trusted(sim, words)
            if host_info and not _auto:
                self.reset_infos = {"GEBV": gebv_t.cpu().numpy()}
            elif host_info:  # copied to the host only if somebody reads them (its own copy: the set is redrawn in two episodes)
                owned, ev = gebv_t.clone(), torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))

                def fetch(buf=owned, ev=ev):
                    ev.synchronize()
                    return buf.cpu().numpy()

                self.reset_infos = _LazyInfos(fetch)
            else:
                self.reset_infos = {"GEBV": gebv_t}
            self._prefetch_next(germ, germ_gebv_ptr, E, n, T)
            return self.populations, self.reset_infos

        ring = self._obs_slots(n, T)
        if ring is not None:
            cur = self.populations
            slot = ring.take(cur.words.data_ptr() if isinstance(cur, PackedPopulation) else 0)
            pop, words_ptr = PackedPopulation._trusted(sim, ring.words[slot]), ring.pop_ptr[slot]
        else:
            words = sim._empty_words(E, n)
            pop, words_ptr = PackedPopulation._trusted(sim, words), words.data_ptr()
        lazy = None
        if host_info and _auto:
            # autoreset inside a step: infos stay on the GPU (their own buffer) until somebody reads env.reset_infos
            lazy = torch.empty((E, n, T), dtype=torch.float32, device=self.device)  # owned by the infos object
            gebv_dev_ptr, gebv_host_ptr = lazy.data_ptr(), None
        elif host_info:
            io = self._host_io((E, n, 2), T)
            gebv_dev_ptr, gebv_host_ptr = io["gebv_dev"], io["gebv_pin"]
        elif ring is not None:
            gebv_dev_ptr, gebv_host_ptr = ring.gebv_ptr[slot], None
        else:
            gebv_t = torch.empty((E, n, T), dtype=torch.float32, device=self.device)
            gebv_dev_ptr, gebv_host_ptr = gebv_t.data_ptr(), None
        # env g draws permutation(keys[1 + g], N)[:n] with keys = split(random_key, total + 1); one C call does
        # the draw, the gather from the germplasm and the reset infos
        rc = self._vec_reset_fn(sim._engine, germ.data_ptr(), germ.shape[0], _lib.nptr(key), total, begin, E, n, sim._layout_id,
                                idx.data_ptr(), words_ptr, gebv_dev_ptr, gebv_host_ptr, germ_gebv_ptr,
                                self._raw_stream(self._dev_index))
        if rc:
            _lib.check(rc)
        if lazy is not None:
            done_ev = torch.cuda.Event()
            done_ev.record(torch.cuda.current_stream(self.device))

            def fetch(buf=lazy, ev=done_ev):
                ev.synchronize()
                return buf.cpu().numpy()

            infos = _LazyInfos(fetch)
        elif host_info:
            infos = {"GEBV": io["gebv_np"].copy()}
        elif ring is not None:
            infos = ring.infos[slot]
        else:
            infos = {"GEBV": gebv_t}
        self.random_key = _lib.key_split_at(self.random_key, 0, total + 1, sim.rng_layout)
        self._reset_indices = idx
        self.populations = self._own = pop
        self.reset_infos = infos
        if self.reset_prefetch and self.autoreset and germ_gebv_ptr is not None:
            self._prefetch_next(germ, germ_gebv_ptr, E, n, T)
        return self.populations, self.reset_infos

    def _prefetch_next(self, germ, germ_gebv_ptr, E, n, T):
        """Draw the reset the NEXT autoreset will ask for (same `random_key`, same n) on the library's side stream."""
        sim = self.simulator
        begin, total = self.env_shard
        pre = self._pre
        if pre is None or pre["n"] != n or pre["E"] != E:
            bufs = [(sim._empty_words(E, n), torch.empty((E, n, T), dtype=torch.float32, device=self.device),
                     torch.empty((E, n), dtype=torch.int32, device=self.device)) for _ in range(2)]
            pre = self._pre = {"n": n, "E": E, "bufs": bufs, "slot": 0, "pending": False, "key": None}
        pre["slot"] ^= 1  # (the other set holds the populations adopted last: still the current episode's first parents)
        words, gebv_t, idx = pre["bufs"][pre["slot"]]
        key = np.array(self.random_key, dtype=np.uint32, copy=True)
        rc = self._prefetch_fn(sim._engine, germ.data_ptr(), germ.shape[0], _lib.nptr(key), total, begin, E, n, sim._layout_id,
                               idx.data_ptr(), words.data_ptr(), gebv_t.data_ptr(), germ_gebv_ptr, self._raw_stream(self._dev_index))
        if rc:
            _lib.check(rc)
        pre["key"], pre["pending"] = key, True

    def join(self):
        """Order the current stream behind the work this env's engine runs on its own side stream (the crossover masks
        of the following steps, the prefetched reset): `bg_engine_join`.  For timing harnesses."""
        rc = _lib.load().bg_engine_join(self.simulator._engine, self._raw_stream(self._dev_index))
        if rc:
            _lib.check(rc)

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
