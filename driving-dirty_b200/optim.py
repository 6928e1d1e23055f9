"""Fused Adam for the scene pipeline (reference: ``torch.optim.Adam(self.parameters(), lr)`` at
roadmap_bce_v2.py:155 / autoencoder.py:120, stepped by Lightning after ddp's gradient all-reduce).

One process per GPU.  With world size 1 every tensor is updated by one ``dd_adam_step`` launch.
With world size N the parameters are *sharded across ranks for the update only*: gradient and
weight replicas live in one symmetric-memory buffer per rank
(``torch.distributed._symmetric_memory``: every replica is peer-mapped into every process, plus an
NVLink multicast mapping where the fabric has one), and ``dd_adam_step_sharded`` does reduce + Adam
+ all-gather in ONE kernel (csrc/adam.cu): 1/N of the optimizer state and of the optimizer's HBM
traffic per GPU, and no NCCL kernel in the step.

* The wide FC weights (>= ``shard_min_numel`` elements; > 99.9 % of the bytes) each own a region;
  ``ops.linear``'s weight-gradient kernel writes straight into the gradient replica (``p.grad`` is a
  persistent view that every backward pass OVERWRITES; ``zero_grad`` leaves it in place).  With
  ``overlap_backward=True`` their update is launched from the backward pass itself, as soon as the
  gradient is written, on a side stream with one small CTA per SM that co-resides with the
  persistent conv kernels of the remaining backward pass; ``step()`` joins it.
* All other tensors share one flat region: their gradients are gathered into it by one multi-tensor
  copy and updated by one sharded launch in ``step()``.

``state_dict()`` / ``load_state_dict()`` speak torch.optim.Adam's layout (per-parameter ``step`` / ``exp_avg`` /
``exp_avg_sq`` of the parameter's shape) at every world size: the moment shards live in a second symmetric buffer, so
the rank that writes a checkpoint (rank 0 under Lightning) assembles the full moments from its peers' shards without a
collective, and every rank re-slices its own shard on load -- a checkpoint moves between world sizes and to / from
torch.optim.Adam.

Semantics are torch.optim.Adam's (amsgrad off, L2 weight decay); ``param_groups[i]['lr']`` is read
every step, so ``ReduceLROnPlateau`` (roadmap_bce_v2.py:156) works unchanged; tensors whose
gradient is ``None`` are skipped (a partly frozen flat bucket is updated per tensor on every replica).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, stream_ptr

_ALIGN = 64          # elements: every tensor of the flat region starts on a 256-byte boundary


def shard_bounds(numel: int, world: int, rank: int):
    """Elements [lo, hi) of a flat tensor that ``rank`` updates: equal shards of 4-element (16 B)
    granules, the last rank takes the remainder granules.  numel must be a multiple of 4."""
    granules = numel // 4
    per = granules // world
    lo = rank * per
    hi = granules if rank == world - 1 else lo + per
    return 4 * lo, 4 * hi


def flat_layout(numels, align: int = _ALIGN):
    """Offsets of tensors packed into one flat region (each start aligned) and the region's length."""
    offs, off = [], 0
    for n in numels:
        offs.append(off)
        off += -(-n // align) * align
    return offs, off


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None,
                 shard_min_numel: int = 1 << 20, multicast: bool | None = None, broadcast_init: bool = True,
                 overlap_backward: bool = False, fuse_into_backward: bool = False, defer_join: bool = True):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) > 1 and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            raise NotImplementedError("FusedAdam: the sharded form supports one parameter group")
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self._regions = {}          # id(param) -> region dict of a wide tensor
        self._flat = None           # region dict of the flat bucket
        self._symm = None
        self._overlap = bool(overlap_backward)
        self._defer_join = bool(defer_join)
        self._side = None           # stream of the updates launched from the backward pass
        self._written = set()       # ids of wide params whose gradient replica was written since the last step()
        self._launched = set()      # ... and whose update is already running on the side stream
        self._mom, self._mom_of = None, None   # symmetric buffer of this rank's moment shards, accessor for the peers'
        self._flat_replicated = False   # the flat bucket fell back to per-tensor replicated updates (partly frozen model)
        self._lib = _lib.load()     # raises when the CUDA library is missing: no fallback
        for g in self.param_groups:
            for p in g["params"]:
                self._check_param(p)
        self._fused = {}            # id(param) -> True once its step ran inside the backward pass (world size 1)
        if self.world > 1:
            self._setup_sharding(shard_min_numel, multicast, broadcast_init)
        elif fuse_into_backward:
            self._setup_fused_backward(shard_min_numel)

    # The two places where the device is touched outside `_launch_sharded`; tests/test_optim.py overrides them (and the
    # launch) to run the sharding logic itself under gloo on CPU tensors.
    def _check_param(self, p):
        if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
            raise RuntimeError("FusedAdam: parameters must be contiguous fp32 CUDA tensors")

    def _alloc_symmetric(self, numel, device, multicast):
        """One symmetric-memory buffer of `numel` floats per rank -> (buffer, handle with .barrier(channel=),
        per-rank device pointers, multicast address or 0)."""
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(numel, dtype=torch.float32, device=device)
        grp = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(buf, grp)
        # per NVLink direction and GPU the multicast form moves n(1 + 1/N) bytes (the switch also loops the own replica
        # back), peer loads / stores 2n(N-1)/N.  Alone, the multicast update is the shorter kernel from N = 6 on (2.0 vs
        # 3.0 ms at N = 8), but it runs BESIDE the conv backward pass, and multimem traffic slows those HBM-bound kernels far
        # more than peer loads do (N = 8, profiles/r2_step_timeline_n8_*.txt: input-gradient conv 1.36 vs 0.82 ms, c3
        # weight gradient 0.69 vs 0.28): 6.68 vs 6.46 ms per step at N = 8, 6.24 vs 5.66 at N = 2.  Peer loads by default.
        want_mc = False if multicast is None else bool(multicast)
        mc = int(hdl.multicast_ptr) if (want_mc and hdl.has_multicast_support) else 0
        if multicast is True and mc == 0:
            raise RuntimeError("FusedAdam: multicast requested but the symmetric buffer has no multicast mapping")
        return buf, hdl, [int(x) for x in hdl.buffer_ptrs], mc

    def _alloc_moments(self, numel, device):
        """Symmetric buffer for the moment shards -> (local tensor, fn(rank) -> that rank's tensor)."""
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(numel, dtype=torch.float32, device=device)
        grp = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(buf, grp)
        buf.zero_()
        return buf, (lambda rk: buf if rk == self.rank else hdl.get_buffer(rk, (numel,), torch.float32))

    @staticmethod
    def _alias(view):
        """A tensor over the same device memory as `view` but with a storage of its own, exactly as long as the tensor:
        torch.save of a parameter that is a plain view would serialise the whole [gradients | weights] buffer."""
        if view.device.type != "cuda":
            return view

        class _Mem:
            __cuda_array_interface__ = {"shape": tuple(view.shape), "typestr": "<f4", "data": (view.data_ptr(), False),
                                        "version": 3, "strides": None}

        return torch.as_tensor(_Mem(), device=view.device)

    # ------------------------------------------------------------------------------------------
    def _setup_sharding(self, shard_min_numel, multicast, broadcast_init):
        params = [p for g in self.param_groups for p in g["params"]]
        if broadcast_init:          # replicas must start identical (Lightning ddp broadcasts rank 0's weights)
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            for p in params:
                dist.broadcast(p.data, src=src, group=self.group)
        wide = [p for p in params if p.numel() >= shard_min_numel and p.numel() % 4 == 0]
        small = [p for p in params if not any(p is q for q in wide)]
        offs, flat_off = flat_layout([p.numel() for p in wide])
        small_offs, small_total = flat_layout([p.numel() for p in small])
        total = flat_off + small_total
        if total == 0:
            return
        dev = params[0].device
        buf, hdl, ptrs, mc = self._alloc_symmetric(2 * total, dev, multicast)      # [gradients | weights]
        buf.zero_()

        def adopt(p, off):
            n = p.numel()
            w = self._alias(buf[total + off: total + off + n].view_as(p))
            w.copy_(p.data)
            p.data = w                                   # the module now computes from the symmetric replica
            return buf[off: off + n].view_as(p)

        for p, off in zip(wide, offs):
            gview = adopt(p, off)
            p.grad = gview
            p._dd_grad_buffer = gview                    # ops.linear's weight-gradient kernel writes here ...
            p._dd_grad_ready = (lambda q: (lambda: self._on_grad_written(q)))(p)     # ... and then calls this
            lo, hi = shard_bounds(p.numel(), self.world, self.rank)
            self._regions[id(p)] = dict(off=off, n=p.numel(), lo=lo, hi=hi, channel=1 + len(self._regions), key=p)
        if small:
            views = [adopt(p, flat_off + o) for p, o in zip(small, small_offs)]
            lo, hi = shard_bounds(small_total, self.world, self.rank)
            self._flat = dict(off=flat_off, n=small_total, lo=lo, hi=hi, channel=0, params=small, grad_views=views, key="flat",
                              offs=small_offs)
        # moment shards: [exp_avg | exp_avg_sq] per region, capacity = the longest shard (the last rank's), in a symmetric
        # buffer of their own so that state_dict() can read the peers' shards
        moff = 0
        for r in list(self._regions.values()) + ([self._flat] if self._flat is not None else []):
            lo_l, hi_l = shard_bounds(r["n"], self.world, self.world - 1)
            cap = -(-max(hi_l - lo_l, r["hi"] - r["lo"], 4) // _ALIGN) * _ALIGN
            r["moff"], r["mcap"] = moff, cap
            moff += 2 * cap
        self._mom, self._mom_of = self._alloc_moments(moff, dev)
        self._symm = dict(buf=buf, hdl=hdl, total=total, ptrs=ptrs, mc=mc)
        if self._overlap:
            self._side = torch.cuda.Stream(device=dev)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        hdl.barrier()

    def _setup_fused_backward(self, min_numel):
        """World size 1: the wide 2-D weights (the FC layers: > 99.9 % of the parameters) get their Adam step INSIDE the
        backward pass, in the epilogue of ops.linear's weight-gradient kernel (dd_linear_wgrad_adam): the gradient tile is
        consumed where it is produced and never written to HBM (24 bytes per parameter instead of 32, and no gradient
        buffers: 1.3 GB less for RoadMapBCE).  Measured at the bench shape the step time is the same as the unfused pair of
        kernels (5.94 ms both: the epilogue's scattered 128-byte accesses to three arrays reach 4.4 TB/s against the
        streaming Adam kernel's 6.2), so it is opt-in (``fuse_into_backward=True``).  step() then skips them.  Like the sharded form's
        overlap_backward this is for loops that step after every backward (the reference's); ops.linear falls back to a
        stored gradient whenever the fused kernel does not apply (batch > 32, CUDA-core path), and step() handles that."""
        for g in self.param_groups:
            for p in g["params"]:
                if p.dim() == 2 and p.numel() >= min_numel and p.is_cuda:
                    p._dd_fused_update = (lambda q, grp: (lambda dy, x, db: self._fused_step(grp, q, dy, x, db)))(p, g)

    @torch.no_grad()
    def _fused_step(self, g, p, dy, x, db):
        if id(p) in self._fused:
            raise RuntimeError("FusedAdam: two backward passes reached a weight whose Adam step is folded into its backward, "
                               "without a step() in between")
        st = self._state(p, p.numel(), p.device)
        st["step"] += 1
        B, K = x.shape
        call("dd_linear_wgrad_adam", dy.data_ptr(), x.data_ptr(), p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
             db.data_ptr() if db is not None else None, B, p.shape[0], K, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
             float(g["eps"]), float(g["weight_decay"]), st["step"], stream_ptr())
        self._fused[id(p)] = True

    @property
    def uses_multicast(self) -> bool:
        return bool(self._symm and self._symm["mc"])

    # ------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if id(p) in self._regions:       # persistent view of the symmetric buffer, overwritten by every backward
                    continue
                if p.grad is None:
                    continue
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_()
                    p.grad.zero_()

    def _state(self, key, numel, device):
        st = self.state[key]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros(numel, dtype=torch.float32, device=device)
            st["exp_avg_sq"] = torch.zeros(numel, dtype=torch.float32, device=device)
        return st

    def _shard_state(self, r):
        """State of a sharded region: views of this rank's piece of the symmetric moments buffer."""
        st = self.state[r["key"]]
        if not st:
            n = r["hi"] - r["lo"]
            st["step"] = 0
            st["exp_avg"] = self._mom[r["moff"]: r["moff"] + n]
            st["exp_avg_sq"] = self._mom[r["moff"] + r["mcap"]: r["moff"] + r["mcap"] + n]
        return st

    def _hyper(self):
        g = self.param_groups[0]
        return (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.world == 1:
            for g in self.param_groups:
                for p in g["params"]:
                    if id(p) in self._fused:              # stepped inside the backward pass
                        continue
                    if p.grad is not None:
                        self._launch_local(g, p, p.grad)
            self._fused.clear()
            return loss
        if self._symm is None:
            return loss
        hdl = self._symm["hdl"]
        todo = [r for key, r in self._regions.items() if key in self._written and key not in self._launched]
        fl = self._flat
        if fl is not None:
            have = [p.grad is not None for p in fl["params"]]
            if any(have) and (self._flat_replicated or not all(have)):
                # A partly frozen bucket (e.g. RoadMapBCE before unfreeze_epoch_no, roadmap_bce_v2.py:127-129):
                # torch.optim.Adam skips tensors without a gradient and keeps a step count per tensor, which one sharded
                # bucket cannot express.  These tensors are 0.1 % of the bytes: average each present gradient with an
                # all-reduce and update every replica locally, per tensor -- and stay in this form from then on, so the
                # per-tensor moments remain the truth.
                if self.state.get("flat"):
                    raise NotImplementedError("FusedAdam (world > 1): the flat bucket was already stepped in its sharded form; "
                                              "freeze parameters before the first step, not after")
                self._flat_replicated = True
                nccl = dist.get_backend(self.group) == "nccl"
                g = self.param_groups[0]
                for p in fl["params"]:
                    if p.grad is None:
                        continue
                    dist.all_reduce(p.grad, op=dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM, group=self.group)
                    if not nccl:
                        p.grad.div_(self.world)
                    self._launch_local(g, p, p.grad)
            elif all(have):
                torch._foreach_copy_(fl["grad_views"], [p.grad for p in fl["params"]])
                todo.append(fl)
        if todo:
            hdl.barrier()                                  # every rank has written its gradient replicas
            for r in todo:
                self._launch_sharded(r, ctas_per_sm=8)
            hdl.barrier()                                  # every replica holds the new weights
        if self._launched:
            # The updates launched from the backward pass run on the side stream and end with a cross-rank barrier (new
            # weights landed in every replica, every rank done reading the gradient replicas).  Their tail (the NVLink
            # exchange of the last wide tensor) need not finish before step() returns -- only before the next READ of that
            # weight: each adopted parameter carries an event that ops.linear waits for in the next forward pass, so the tail
            # overlaps with the next step's conv forward instead of being exposed (N = 8: ~1 ms per step).  synchronize()
            # joins explicitly (checkpointing, evaluation code that reads the parameters directly).
            if self._defer_join:
                ev = torch.cuda.Event()
                ev.record(self._side)
                for key in self._launched:
                    self._regions[key]["key"]._dd_ready_event = ev
                self._pending = ev
            else:
                torch.cuda.current_stream().wait_stream(self._side)
        self._launched.clear()
        self._written.clear()
        return loss

    def synchronize(self):
        """Make the current stream wait for weight updates still running on the side stream (overlap_backward)."""
        ev = getattr(self, "_pending", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self._pending = None

    # ------------------------------------------------------------------------------------------
    # checkpointing: torch.optim.Adam's layout in and out
    # ------------------------------------------------------------------------------------------
    def _full_moments(self, r):
        """(exp_avg, exp_avg_sq) of a sharded region at full length, assembled from every rank's shard."""
        dev = self._mom.device
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)      # this rank's step() ended with a cross-rank barrier: every shard is up to date
        m, v = torch.empty(r["n"], dtype=torch.float32, device=dev), torch.empty(r["n"], dtype=torch.float32, device=dev)
        for rk in range(self.world):
            lo, hi = shard_bounds(r["n"], self.world, rk)
            peer = self._mom_of(rk)
            m[lo:hi].copy_(peer[r["moff"]: r["moff"] + hi - lo])
            v[lo:hi].copy_(peer[r["moff"] + r["mcap"]: r["moff"] + r["mcap"] + hi - lo])
        return m, v

    def state_dict(self):
        self.synchronize()
        index, groups = {}, []
        for g in self.param_groups:
            pg = {k: v for k, v in g.items() if k != "params"}
            pg["params"] = []
            for p in g["params"]:
                index[id(p)] = len(index)
                pg["params"].append(index[id(p)])
            groups.append(pg)
        state = {}

        def put(p, step, m, v):
            state[index[id(p)]] = {"step": torch.tensor(float(step), dtype=torch.float32),
                                   "exp_avg": m.detach().clone().view_as(p), "exp_avg_sq": v.detach().clone().view_as(p)}

        for g in self.param_groups:
            for p in g["params"]:
                st = self.state.get(p)
                if st and id(p) not in self._regions:        # world 1, or the per-tensor (replicated) form
                    put(p, st["step"], st["exp_avg"], st["exp_avg_sq"])
        for key, r in self._regions.items():
            st = self.state.get(r["key"])
            if st and not isinstance(r["key"], str):
                m, v = self._full_moments(r)
                put(r["key"], st["step"], m, v)
        fl = self._flat
        if fl is not None and self.state.get("flat"):
            m, v = self._full_moments(fl)
            for p, o in zip(fl["params"], fl["offs"]):
                put(p, self.state["flat"]["step"], m[o: o + p.numel()], v[o: o + p.numel()])
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, state_dict):
        params = [p for g in self.param_groups for p in g["params"]]
        saved_groups = state_dict["param_groups"]
        ids = [i for g in saved_groups for i in g["params"]]
        if len(ids) != len(params):
            raise ValueError("FusedAdam.load_state_dict: the checkpoint has a different number of parameters")
        for g, sg in zip(self.param_groups, saved_groups):
            for k, v in sg.items():
                if k != "params" and k in g:
                    g[k] = v
        by_param = {id(p): state_dict["state"].get(i, state_dict["state"].get(str(i))) for p, i in zip(params, ids)}

        def unpack(p):
            st = by_param.get(id(p))
            if not st:
                return None
            step = st["step"]
            step = int(round(float(step.item() if torch.is_tensor(step) else step)))
            dev = p.device
            m = st["exp_avg"].detach().to(device=dev, dtype=torch.float32).reshape(-1)
            v = st["exp_avg_sq"].detach().to(device=dev, dtype=torch.float32).reshape(-1)
            if m.numel() != p.numel():
                raise ValueError("FusedAdam.load_state_dict: moment shape does not match the parameter")
            return step, m, v

        self.state.clear()
        for p in params:
            if id(p) in self._regions:
                got = unpack(p)
                if got:
                    r = self._regions[id(p)]
                    st = self._shard_state(r)
                    st["step"] = got[0]
                    st["exp_avg"].copy_(got[1][r["lo"]: r["hi"]])
                    st["exp_avg_sq"].copy_(got[2][r["lo"]: r["hi"]])
        fl = self._flat
        flat_ids = {id(p) for p in fl["params"]} if fl is not None else set()
        if fl is not None:
            got = [unpack(p) for p in fl["params"]]
            steps = {g[0] for g in got if g}
            if all(got) and len(steps) == 1:                 # one step count: the sharded form can carry it
                dev = self._mom.device
                m, v = torch.zeros(fl["n"], device=dev), torch.zeros(fl["n"], device=dev)
                for p, o, g in zip(fl["params"], fl["offs"], got):
                    m[o: o + p.numel()].copy_(g[1])
                    v[o: o + p.numel()].copy_(g[2])
                st = self._shard_state(fl)
                st["step"] = steps.pop()
                st["exp_avg"].copy_(m[fl["lo"]: fl["hi"]])
                st["exp_avg_sq"].copy_(v[fl["lo"]: fl["hi"]])
                self._flat_replicated = False
            elif any(got):                                   # per-tensor step counts (partly frozen history): replicated form
                self._flat_replicated = True
                for p, g in zip(fl["params"], got):
                    if g:
                        self.state[p] = {"step": g[0], "exp_avg": g[1].clone(), "exp_avg_sq": g[2].clone()}
        for p in params:
            if id(p) in self._regions or id(p) in flat_ids:
                continue
            got = unpack(p)
            if got:
                self.state[p] = {"step": got[0], "exp_avg": got[1].clone(), "exp_avg_sq": got[2].clone()}

    def _launch_local(self, g, p, grad):
        """One dd_adam_step launch: the whole tensor, this replica only."""
        if not (grad.is_contiguous() and grad.dtype == torch.float32):
            grad = grad.contiguous().float()
        st = self._state(p, p.numel(), p.device)
        st["step"] += 1
        call("dd_adam_step", p.data_ptr(), grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
             p.numel(), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
             float(g["weight_decay"]), st["step"], 1.0, stream_ptr())

    def _launch_sharded(self, r, ctas_per_sm):
        sy, W = self._symm, self.world
        lo, hi, total = r["lo"], r["hi"], sy["total"]
        st = self._shard_state(r)
        st["step"] += 1
        gptrs = (ctypes.c_void_p * W)(*[b + 4 * r["off"] for b in sy["ptrs"]])
        wptrs = (ctypes.c_void_p * W)(*[b + 4 * (total + r["off"]) for b in sy["ptrs"]])
        mcg = ctypes.c_void_p(sy["mc"] + 4 * r["off"]) if sy["mc"] else None
        mcw = ctypes.c_void_p(sy["mc"] + 4 * (total + r["off"])) if sy["mc"] else None
        lr, b1, b2, eps, wd = self._hyper()
        call("dd_adam_step_sharded", gptrs, wptrs, mcg, mcw, W, self.rank, st["exp_avg"].data_ptr(),
             st["exp_avg_sq"].data_ptr(), lo, hi - lo, lr, b1, b2, eps, wd, st["step"], 1.0 / W, int(ctas_per_sm), stream_ptr())

    @torch.no_grad()
    def _on_grad_written(self, p):
        """Called by ops.linear's backward right after p's gradient replica is written.  overlap_backward: the reduce
        + Adam + all-gather kernel is launched here, on a side stream, one 256-thread CTA per SM beside the persistent
        conv kernels of the rest of the backward pass; step() joins the side stream.  (The weights then change before
        step() is called: for training loops that step after every backward, like the reference's.)"""
        key = id(p)
        if key in self._written:
            raise RuntimeError("FusedAdam: two backward passes wrote the gradient of a sharded parameter without a "
                               "step() in between (its gradient buffer is overwritten, not accumulated)")
        self._written.add(key)
        if not self._overlap:
            return
        r = self._regions[key]
        hdl = self._symm["hdl"]
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            hdl.barrier(channel=r["channel"])
            self._launch_sharded(r, ctas_per_sm=1)
            hdl.barrier(channel=r["channel"])
        self._launched.add(key)


def make_adam(module, lr):
    """What the task modules' ``configure_optimizers`` return for ``torch.optim.Adam(self.parameters(), lr)``
    (roadmap_bce_v2.py:155, autoencoder.py:120, spatial_w_rm.py:189).  hparams.optimizer = "fused" (the default when the
    model sits on a CUDA device): ``FusedAdam`` -- the dd_adam_step kernels, and under torch.distributed the sharded
    reduce + Adam + all-gather exchange, launched from the backward pass; "torch": the reference's optimizer object."""
    hp = getattr(module, "hparams", None)
    kind = getattr(hp, "optimizer", "fused") if hp is not None else "fused"
    params = [p for p in module.parameters()]
    on_cuda = bool(params) and all(p.is_cuda for p in params)
    if kind == "fused" and on_cuda:
        opt = FusedAdam(params, lr=lr, overlap_backward=dist.is_initialized() and dist.get_world_size() > 1)
        # a checkpoint (module.state_dict()) must see the weights of updates that are still in flight on the side stream
        module.register_state_dict_pre_hook(lambda *a, **k: opt.synchronize())
        return opt
    return torch.optim.Adam(params, lr=lr)
