"""Multi-GPU state vector: one process per GPU, the top log2(P) qubits are the rank id.

The reference is single-GPU only (its README lists multi-GPU as future work), so this module has no
reference counterpart; its contract is "same amplitudes as the single-device run".

  * gates whose non-diagonal target is a local qubit run in the fused-pass engine on every shard;
    controls and diagonal gates on global qubits need no data movement (the rank supplies the bit);
  * an uncontrolled X on a global qubit is a rank relabelling, carried as an X frame;
  * any other non-diagonal gate on a global qubit g first swaps g with a local qubit l: rank r and
    r ^ (1 << (g - n_local)) exchange the half of their shards whose bit l differs from their own
    value of g (16 * 2^(n_local - 1) bytes each way: the NVLink roofline of the step).  The swap is
    never undone: the logical -> physical qubit permutation is carried and resolved at read-out.

`plan_circuit` (pure host logic) decides where the swaps go; an *engine* executes the steps.  The
product engine is `CudaShardEngine` (C ABI + CUDA IPC peer memory, or NCCL send/recv through
torch.distributed).  Tests inject their own engine to exercise the planning / exchange choreography
on CPU with the gloo backend; the product never falls back to anything but CUDA.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import byref, c_double, c_int64, c_uint64, c_void_p
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import GATE_DTYPE
from .circuit import Circuit

_DIAGONAL = {2, 4, 5, 6, 7, 10, 12, 14}      # Z S T Sdag Tdag Rz CZ CRZ
_X = 0


def _target_positions(g) -> Tuple[int, ...]:
    """Qubit slots (q0/q1/q2) of a gate record that are non-diagonal targets."""
    t = int(g["type"])
    if t in _DIAGONAL:
        return ()
    if t <= 10:
        return (0,)
    if t == 15:                # SWAP: both are targets
        return (0, 1)
    if t == 16:
        return (2,)
    return (1,)                # CNOT, CRY


def choose_initial_layout(num_qubits: int, n_global: int, gates: np.ndarray) -> List[int]:
    """Logical qubit -> physical position for a circuit that starts from |0...0> (which is the same state under any
    relabelling of the qubits): the qubits that are never a non-diagonal target — uncontrolled X is an index relabel,
    diagonal gates and controls need no data — go to the global positions, where they cost no exchange at all; if there
    are not enough of those, the ones targeted last.  Everything else keeps its relative order."""
    n, nl = num_qubits, num_qubits - n_global
    first_use = [1 << 60] * n
    qcols = ("q0", "q1", "q2")
    for j, g in enumerate(gates):
        if int(g["type"]) == _X:
            continue
        for slot in _target_positions(g):
            q = int(g[qcols[slot]])
            if first_use[q] == 1 << 60:
                first_use[q] = j
    # best candidates for the global positions: latest first non-diagonal use, ties -> highest qubit (identity if possible)
    order = sorted(range(n), key=lambda q: (first_use[q], q), reverse=True)
    glob = sorted(order[:n_global])
    perm = [0] * n
    pos = 0
    for q in range(n):
        if q not in glob:
            perm[q] = pos
            pos += 1
    for i, q in enumerate(glob):
        perm[q] = nl + i
    return perm


@dataclass
class Step:
    kind: str                              # "gates" | "swap"
    gates: Optional[np.ndarray] = None     # physical-qubit gate records
    global_qubit: int = -1
    local_qubit: int = -1


@dataclass
class Plan:
    num_qubits: int
    n_global: int
    steps: List[Step] = field(default_factory=list)
    perm: List[int] = field(default_factory=list)      # logical qubit -> physical position after the plan
    n_gates: int = 0

    @property
    def n_swaps(self) -> int:
        return sum(1 for s in self.steps if s.kind == "swap")


def plan_circuit(num_qubits: int, n_global: int, gates: np.ndarray, perm: Optional[Sequence[int]] = None) -> Plan:
    """Split a circuit into local segments separated by global<->local qubit swaps.

    `perm[q]` is the physical bit position currently holding logical qubit q.  The swap partner is the
    local position whose next use as a non-diagonal target lies farthest in the future (Belady); among equals, one
    that is unlikely to be a tile qubit of the preceding pass (so the exchange can be fused into it), then the highest
    position (long contiguous runs in the exchange).
    """
    n, nl = num_qubits, num_qubits - n_global
    perm = list(range(n)) if perm is None else list(perm)
    plan = Plan(n, n_global, n_gates=len(gates))
    cur: List[tuple] = []

    def flush():
        if cur:
            plan.steps.append(Step("gates", gates=np.array(cur, dtype=GATE_DTYPE)))
            cur.clear()

    qcols = ("q0", "q1", "q2")
    for idx, g in enumerate(gates):
        t = int(g["type"])
        if t != _X:   # an uncontrolled X on a global qubit is a frame toggle, handled by the compiler
            for slot in _target_positions(g):
                lq = int(g[qcols[slot]])
                if perm[lq] >= nl:
                    # tile qubits of the segment's last pass, roughly: the low contiguous run plus the most recent
                    # non-diagonal targets.  A victim outside them lets the exchange ride on that pass's store.
                    recent = set(range(min(5, nl)))
                    for rec in reversed(cur):
                        if len(recent) >= 12:
                            break
                        if rec[0] == _X:
                            continue
                        rg = {"type": rec[0]}
                        for slot in _target_positions(rg):
                            recent.add(rec[1 + slot])
                    flush()
                    # choose the local position to evict
                    next_use = {p: 1 << 60 for p in range(nl)}
                    inv = {perm[q]: q for q in range(n)}
                    busy = {perm[int(g[c])] for c in qcols if int(g[c]) >= 0}
                    for j in range(idx, len(gates)):
                        h = gates[j]
                        for s2 in _target_positions(h):
                            p = perm[int(h[qcols[s2]])]
                            if p < nl and next_use[p] == 1 << 60:
                                next_use[p] = j
                    cand = [p for p in range(nl) if p not in busy]
                    victim = max(cand, key=lambda p: (next_use[p], p not in recent, p))
                    gpos = perm[lq]
                    plan.steps.append(Step("swap", global_qubit=gpos, local_qubit=victim))
                    other = inv[victim]
                    perm[lq], perm[other] = victim, gpos
        rec = (t, perm[int(g["q0"])] if int(g["q0"]) >= 0 else -1, perm[int(g["q1"])] if int(g["q1"]) >= 0 else -1,
               perm[int(g["q2"])] if int(g["q2"]) >= 0 else -1, float(g["param"]))
        cur.append(rec)
    flush()
    plan.perm = perm
    return plan


# ---------------------------------------------------------------------------------------------------


class CudaShardEngine:
    """Executes plan steps on this rank's GPU shard through the C ABI."""

    def __init__(self, num_qubits: int, n_global: int, rank: int, world: int, exchange: str = "auto"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.n, self.ng, self.nl = num_qubits, n_global, num_qubits - n_global
        self.rank, self.world = rank, world
        self.state = torch.empty(1 << self.nl, dtype=torch.complex128, device="cuda")
        self._h = c_void_p()
        _lib.check(_lib.lib().qsim_shard_create(self.n, self.ng, rank, c_void_p(self.state.data_ptr()), byref(self._h)))
        self.stream = torch.cuda.current_stream()
        _lib.check(_lib.lib().qsim_sim_set_stream(self._h, c_void_p(self.stream.cuda_stream)))
        self._flag = torch.zeros(1, device="cuda")
        self._peer_ptr = {}          # peer rank -> [pointer to its buffer 0, pointer to its buffer 1]
        self._peer_base = []
        self.exchange = exchange
        self._bounce = None
        # Second shard buffer for the fused exchange (the last pass before a swap stores out of place, partly into the
        # partner's second buffer).  Every rank flips buffers at the same steps, so `_cur` is the same number everywhere.
        self._bufs = [self.state]
        self._cur = 0
        self.fused_exchanges = 0
        if world > 1 and exchange in ("auto", "p2p"):
            try:
                free, _total = torch.cuda.mem_get_info()
                shard_bytes = 16 << self.nl
                want_alt = os.environ.get("QSIM_NO_FUSED_EXCHANGE") is None and free > shard_bytes + (4 << 30)
                flags = [None] * world
                dist.all_gather_object(flags, bool(want_alt))
                if all(flags):
                    self._bufs.append(torch.empty(1 << self.nl, dtype=torch.complex128, device="cuda"))
                self._open_peers()
                self.exchange = "p2p"
            except Exception as exc:
                if exchange == "p2p":
                    raise
                import warnings
                warnings.warn(f"qsim_b200 rank {rank}: CUDA-IPC peer memory unavailable ({type(exc).__name__}: {exc}); "
                              "global-qubit swaps fall back to NCCL send/recv through bounce buffers (about a third of the "
                              "peer-to-peer bandwidth)", RuntimeWarning)
                self.exchange = "nccl"
                self._bufs = [self.state]
        elif world > 1:
            self.exchange = "nccl"

    # -- peer memory ----------------------------------------------------------------------------
    def _open_peers(self):
        mine = []
        for t in self._bufs:
            handle = (ctypes.c_ubyte * 64)()
            off = c_uint64()
            _lib.check(_lib.lib().qsim_ipc_get_handle(c_void_p(t.data_ptr()), handle, byref(off)))
            mine.append((bytes(handle), int(off.value)))
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, mine)
        for b in range(self.ng):
            peer = self.rank ^ (1 << b)
            ptrs = []
            for hbytes, poff in everyone[peer]:
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(hbytes)
                base = c_void_p()
                _lib.check(_lib.lib().qsim_ipc_open_handle(buf, byref(base)))
                self._peer_base.append(base)
                ptrs.append(base.value + poff)
            self._peer_ptr[peer] = ptrs

    def device_barrier(self):
        """Stream-ordered barrier across ranks: nobody's later kernels start before everybody's earlier ones ended."""
        if self.world > 1:
            self.dist.all_reduce(self._flag)

    # -- steps ----------------------------------------------------------------------------------
    def compile_gates(self, gates: np.ndarray, initial_xor: int):
        h = c_void_p()
        g = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
        _lib.check(_lib.lib().qsim_program_compile_ex(self.n, self.ng, _lib.gates_ptr(g) if len(g) else None, len(g),
                                                      c_uint64(initial_xor), byref(h)))
        info = (c_int64 * 8)()
        _lib.check(_lib.lib().qsim_program_info(h, info))
        return h, {"passes": info[0], "ops": info[1], "global_xor": int(info[6])}

    def run_program(self, handle):
        _lib.check(_lib.lib().qsim_sim_execute(self._h, handle))

    def free_program(self, handle):
        _lib.lib().qsim_program_destroy(handle)

    def swap(self, global_qubit: int, local_qubit: int):
        peer = self.rank ^ (1 << (global_qubit - self.nl))
        _lib.lib().qsim_sim_device_ptr(self._h)   # a lazily reset shard is written out now: the peer is about to read it
        self.device_barrier()
        if self.exchange == "p2p":
            _lib.check(_lib.lib().qsim_shard_swap_p2p(self._h, c_void_p(self._peer_ptr[peer][self._cur]), global_qubit,
                                                      local_qubit))
        else:
            self._swap_nccl(peer, global_qubit, local_qubit)
        self.device_barrier()

    def run_program_then_swap(self, handle, global_qubit: int, local_qubit: int) -> bool:
        """The program followed by the swap as ONE step: the program's last pass stores out of place, the half that
        leaves going straight into the partner's other buffer over NVLink (qsim_shard_execute_exchange).  Returns
        False, having done nothing, when that is not possible here (the caller then runs the two steps)."""
        if self.exchange != "p2p" or len(self._bufs) < 2:
            return False
        mask = c_uint64()
        _lib.check(_lib.lib().qsim_program_last_tile_mask(handle, byref(mask)))
        if mask.value == 0 or (mask.value >> local_qubit) & 1:
            return False
        peer = self.rank ^ (1 << (global_qubit - self.nl))
        alt = 1 - self._cur
        if os.environ.get("QSIM_EXCHANGE_PREBARRIER"):
            self.device_barrier()
        _lib.check(_lib.lib().qsim_shard_execute_exchange(self._h, handle, c_void_p(self._bufs[alt].data_ptr()),
                                                         c_void_p(self._peer_ptr[peer][alt]), global_qubit, local_qubit))
        self._cur = alt
        self.state = self._bufs[alt]
        self.fused_exchanges += 1
        self.device_barrier()
        return True

    def _swap_nccl(self, peer: int, global_qubit: int, local_qubit: int):
        torch, dist = self.torch, self.dist
        my_bit = (self.rank >> (global_qubit - self.nl)) & 1
        pairs = 1 << (self.nl - 1)
        chunk_amps = min(pairs, 1 << 24)                      # 256 MiB bounce buffers
        n_chunks = pairs // chunk_amps
        if self._bounce is None or self._bounce[0].numel() < chunk_amps:
            self._bounce = (torch.empty(chunk_amps, dtype=torch.complex128, device="cuda"),
                            torch.empty(chunk_amps, dtype=torch.complex128, device="cuda"))
        send, recv = self._bounce
        for c in range(n_chunks):
            # the half that leaves has bit `local_qubit` != my value of the global qubit
            _lib.check(_lib.lib().qsim_shard_pack_half(self._h, local_qubit, my_bit, c, n_chunks, c_void_p(send.data_ptr())))
            ops = [dist.P2POp(dist.isend, send[:chunk_amps], peer), dist.P2POp(dist.irecv, recv[:chunk_amps], peer)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            _lib.check(_lib.lib().qsim_shard_unpack_half(self._h, local_qubit, my_bit, c, n_chunks, c_void_p(recv.data_ptr())))

    # -- state ------------------------------------------------------------------------------------
    def reset(self):
        _lib.check(_lib.lib().qsim_sim_reset(self._h))

    def synchronize(self):
        _lib.check(_lib.lib().qsim_sim_synchronize(self._h))

    def local_state(self) -> np.ndarray:
        out = np.empty(1 << self.nl, np.complex128)
        _lib.check(_lib.lib().qsim_sim_get_state(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def set_local_state(self, amps: np.ndarray):
        a = np.ascontiguousarray(amps, dtype=np.complex128)
        _lib.check(_lib.lib().qsim_sim_set_state(self._h, a.ctypes.data_as(c_void_p)))

    def partial_probability(self, bit: int = -1) -> float:
        v = c_double()
        _lib.check(_lib.lib().qsim_shard_partial_probability(self._h, bit, byref(v)))
        return v.value

    def collapse(self, bit: int, outcome: int, scale: float):
        """Zero the amplitudes whose index bit `bit` != outcome and scale the rest (bit < 0: scale everything)."""
        _lib.check(_lib.lib().qsim_shard_collapse(self._h, int(bit), int(outcome), c_double(scale)))

    def shard_sample(self, c_init: float, first: bool, uniforms: np.ndarray):
        u = np.ascontiguousarray(uniforms, np.float64)
        out = np.empty(len(u), np.int64)
        c_end = c_double()
        _lib.check(_lib.lib().qsim_shard_sample(self._h, c_init, int(first), u.ctypes.data_as(c_void_p), len(u),
                                                out.ctypes.data_as(c_void_p), byref(c_end)))
        return out, c_end.value

    def marginal(self, local_bits) -> np.ndarray:
        qs = np.ascontiguousarray(local_bits, dtype=np.int32)
        out = np.empty(1 << len(qs), np.float64)
        _lib.check(_lib.lib().qsim_sim_marginal(self._h, qs.ctypes.data_as(c_void_p), len(qs), out.ctypes.data_as(c_void_p)))
        return out

    def cdf_prepare(self) -> float:
        v = c_double()
        _lib.check(_lib.lib().qsim_shard_cdf_prepare(self._h, byref(v)))
        return v.value

    def cdf_classify(self, approx_c_init: float):
        _lib.check(_lib.lib().qsim_shard_cdf_classify(self._h, c_double(approx_c_init)))

    def launch_count(self) -> int:
        return int(_lib.lib().qsim_sim_launch_count(self._h))

    def set_timing(self, on: bool):
        _lib.check(_lib.lib().qsim_sim_set_timing(self._h, int(bool(on))))

    def pass_time_ms(self):
        t, n = c_double(), c_int64()
        _lib.check(_lib.lib().qsim_sim_pass_time_ms(self._h, byref(t), byref(n)))
        return t.value, n.value

    def close(self):
        if self._h and self._h.value:
            _lib.lib().qsim_sim_destroy(self._h)
            self._h = c_void_p()
        for b in self._peer_base:
            _lib.lib().qsim_ipc_close_handle(b)
        self._peer_base = []
        # drop the device buffers (peers have closed their mappings by the time they allocate again: close() is collective
        # in practice, every rank calls it at the same point)
        self.state = None
        self._bufs = []
        self._bounce = None


@dataclass
class CompiledPlan:
    plan: Plan
    programs: list                 # per step: program handle or None
    frame_after: int               # X frame (over physical positions) once the plan has run
    n_passes: int
    n_ops: int
    n_swaps: int
    perm_before: List[int] = field(default_factory=list)   # the layout and frame the plan was compiled against
    frame_before: int = 0
    from_pristine: bool = False    # compiled for |0...0> with a freely chosen layout: valid on any pristine state


class ShardedSimulator:
    """Simulator over a state sharded across torch.distributed ranks (one GPU each)."""

    def __init__(self, num_qubits: int, exchange: str = "auto", engine=None, rank: Optional[int] = None,
                 world: Optional[int] = None):
        if engine is None or rank is None:
            import torch.distributed as dist
            world = dist.get_world_size() if dist.is_initialized() else 1
            rank = dist.get_rank() if dist.is_initialized() else 0
        assert world & (world - 1) == 0, "world size must be a power of two"
        self.n = int(num_qubits)
        self.ng = world.bit_length() - 1
        self.nl = self.n - self.ng
        self.rank, self.world = rank, world
        self.engine = engine if engine is not None else CudaShardEngine(self.n, self.ng, rank, world, exchange)
        self.perm = list(range(self.n))      # logical qubit -> physical position
        self.frame = 0                       # pending X mask over physical positions (global bits only, between runs)
        self._pristine = True                # the state is |0...0>: the qubit layout is still free to choose
        self._order_preserving = True        # stored index order == logical index order on the support (see sample)

    @property
    def local(self):
        return self.engine

    # -- execution ----------------------------------------------------------------------------
    def reset(self):
        self.engine.reset()
        self.perm = list(range(self.n))
        self.frame = 0
        self._pristine = True
        self._order_preserving = True

    def set_local_state(self, amps: np.ndarray):
        """Overwrite this rank's shard (stored layout); the qubit layout is fixed from here on."""
        self.engine.set_local_state(amps)
        self._pristine = False

    def compile(self, circuit: Circuit) -> CompiledPlan:
        """Plan + compile against the CURRENT qubit permutation and X frame."""
        if circuit.get_num_qubits() != self.n:
            raise _lib.InvalidArgument("Circuit qubit count doesn't match simulator")
        start_perm = list(self.perm)
        chose = False
        if self._pristine and self.ng > 0 and os.environ.get("QSIM_NO_LAYOUT") is None and not getattr(self, "_identity_only", False):
            start_perm = choose_initial_layout(self.n, self.ng, circuit.gates)   # carried by the plan, applied by execute()
            chose = True
        plan = plan_circuit(self.n, self.ng, circuit.gates, start_perm)
        frame = self.frame
        frame_before = frame
        programs, n_passes, n_ops = [], 0, 0
        for st in plan.steps:
            if st.kind == "gates":
                h, info = self.engine.compile_gates(st.gates, frame)
                programs.append(h)
                n_passes += info["passes"]
                n_ops += info["ops"]
                frame = info["global_xor"] << self.nl       # local part was applied by the program itself
            else:
                programs.append(None)
                g, l = st.global_qubit, st.local_qubit      # the pending X (if any) travels with the qubit
                bg, bl = (frame >> g) & 1, (frame >> l) & 1
                frame = (frame & ~((1 << g) | (1 << l))) | (bl << g) | (bg << l)
        return CompiledPlan(plan, programs, frame, n_passes, n_ops, plan.n_swaps, perm_before=start_perm,
                            frame_before=frame_before, from_pristine=chose)

    def execute(self, cp: CompiledPlan):
        """Runs a compiled plan.  A plan is only valid for the qubit layout and X frame it was compiled against: on the
        state it was compiled for (same permutation and frame) or, for a plan compiled from |0...0> with its own layout,
        on any pristine state.  Re-executing a plan that ends in a different layout than it starts from needs a reset()
        (or a re-compile) in between and is refused here."""
        same_layout = list(cp.perm_before) == list(self.perm) and cp.frame_before == self.frame
        if not same_layout and not (cp.from_pristine and self._pristine and cp.frame_before == self.frame):
            raise _lib.InvalidArgument("plan compiled against a different qubit layout / X frame than the state now has "
                                       "(a plan that changes the layout cannot be executed twice in a row): reset(), or "
                                       "compile it against the current state (compile / compile_sequence)")
        steps, progs = cp.plan.steps, cp.programs
        fuse = getattr(self.engine, "run_program_then_swap", None)
        i = 0
        while i < len(steps):
            st = steps[i]
            if st.kind == "gates":
                nxt = steps[i + 1] if i + 1 < len(steps) else None
                if fuse is not None and nxt is not None and nxt.kind == "swap" and \
                        fuse(progs[i], nxt.global_qubit, nxt.local_qubit):
                    i += 2          # the program's last pass carried the exchange
                    continue
                self.engine.run_program(progs[i])
            else:
                self.engine.swap(st.global_qubit, st.local_qubit)
            i += 1
        self.perm = list(cp.plan.perm)
        self.frame = cp.frame_after
        # a layout chosen for |0...0> keeps the reference's index order on the support as long as no exchange has happened
        # (the parked qubits have one value in every non-zero amplitude and the others keep their relative order)
        self._order_preserving = (self._order_preserving if not self._pristine else True) and cp.n_swaps == 0
        self._pristine = False

    def identity_layout_only(self, on: bool = True):
        """Never choose a layout for |0...0> (benchmarks of the exchange path)."""
        self._identity_only = bool(on)

    def relabel_identity(self):
        """Declares the stored layout to be the identity WITHOUT moving data (a relabelling of the logical qubits)."""
        self.perm = list(range(self.n))
        self.frame = 0
        self._pristine = False
        self._order_preserving = True

    @property
    def has_second_buffer(self) -> bool: return len(getattr(self.engine, "_bufs", [])) > 1
    @property
    def fused_exchanges(self) -> int: return int(getattr(self.engine, "fused_exchanges", 0))
    @property
    def exchange(self) -> str: return getattr(self.engine, "exchange", "none")
    def swap(self, g: int, l: int): self.engine.swap(g, l)

    def compile_sequence(self, circuit: Circuit, k: int) -> List[CompiledPlan]:
        """Plans for k consecutive runs of `circuit` starting from the current state: run i is compiled against the layout
        and X frame run i-1 leaves behind (plans are shared when the layout repeats).  Nothing is executed."""
        saved = (list(self.perm), self.frame, self._pristine)
        cache, out = {}, []
        try:
            for _ in range(k):
                key = (tuple(self.perm), self.frame, self._pristine)
                cp = cache.get(key)
                if cp is None:
                    cp = cache[key] = self.compile(circuit)
                out.append(cp)
                self.perm, self.frame, self._pristine = list(cp.plan.perm), cp.frame_after, False
        finally:
            self.perm, self.frame, self._pristine = saved
        return out

    def release(self, cp: CompiledPlan):
        for h in cp.programs:
            if h is not None:
                self.engine.free_program(h)
        cp.programs = []

    def run(self, circuit: Circuit):
        cp = self.compile(circuit)
        try:
            self.execute(cp)
        finally:
            self.release(cp)

    def synchronize(self):
        self.engine.synchronize()

    # -- index bookkeeping ----------------------------------------------------------------------
    def _logical_rank_order(self) -> List[int]:
        """Physical ranks in the order their shards appear in the *stored* index space with the frame resolved:
        stored rank r holds frame-resolved rank r ^ (frame >> nl)."""
        fx = self.frame >> self.nl
        return [r ^ fx for r in range(self.world)]

    def physical_to_logical_index(self, phys: np.ndarray) -> np.ndarray:
        """Map amplitude indices of the stored layout (frame already resolved) to logical basis-state indices."""
        phys = np.asarray(phys, dtype=np.uint64)
        out = np.zeros_like(phys)
        for q in range(self.n):
            out |= ((phys >> np.uint64(self.perm[q])) & np.uint64(1)) << np.uint64(q)
        return out

    # -- read-out -------------------------------------------------------------------------------
    def _allgather(self, value: float) -> List[float]:
        if self.world == 1:
            return [value]
        return self.engine.allgather_float(value) if hasattr(self.engine, "allgather_float") else self._torch_allgather(value)

    def _torch_allgather(self, value: float) -> List[float]:
        import torch
        import torch.distributed as dist
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    def _allreduce_max(self, arr: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return arr
        if hasattr(self.engine, "allreduce_max"):
            return self.engine.allreduce_max(arr)
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(arr).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def get_total_probability(self) -> float:
        return float(sum(self._allgather(self.engine.partial_probability(-1))))

    def marginal(self, qubits: Sequence[int]) -> np.ndarray:
        """Marginal distribution of up to 12 logical qubits (qubits[i] -> bit i of the outcome), on every rank: each shard
        reduces over its local index bits on the device, the rank bits place the shard's share in the outcome."""
        phys = [self.perm[int(q_)] for q_ in qubits]
        if len(set(phys)) != len(phys):
            raise _lib.InvalidArgument("Duplicate qubit in marginal")
        loc = [(i, p) for i, p in enumerate(phys) if p < self.nl]
        glob = [(i, p) for i, p in enumerate(phys) if p >= self.nl]
        mine = np.ascontiguousarray(self.engine.marginal([p for _, p in loc]), np.float64)
        if self.world > 1:
            if hasattr(self.engine, "allgather_array"):
                shares = [np.asarray(a).view(np.float64) for a in self.engine.allgather_array(mine.view(np.complex128) if len(mine) % 2 == 0 else np.concatenate([mine, [0.0]]).view(np.complex128))]
                shares = [a[:len(mine)] for a in shares]
            else:
                import torch
                import torch.distributed as dist
                t = torch.from_numpy(mine).cuda()
                outs = [torch.empty_like(t) for _ in range(self.world)]
                dist.all_gather(outs, t)
                shares = [o.cpu().numpy() for o in outs]
        else:
            shares = [mine]
        fx = self.frame >> self.nl
        out = np.zeros(1 << len(phys), np.float64)
        sub = np.arange(1 << len(loc))
        spread = np.zeros(len(sub), np.int64)                   # local outcome bits moved to their places
        for j, (i, _) in enumerate(loc):
            spread |= ((sub >> j) & 1) << i
        for r, share in enumerate(shares):
            rank_bits = r ^ fx                                  # stored rank r holds frame-resolved rank r ^ fx
            base = 0
            for i, p in glob:
                base |= ((rank_bits >> (p - self.nl)) & 1) << i
            np.add.at(out, spread | base, share)
        return out

    def restore_identity_layout(self):
        """Moves the amplitudes back to the identity qubit layout (logical qubit q at index bit q): global<->local
        exchanges for the rank bits, then ONE local pass of SWAP gates (pure index permutations: the compiler folds them
        into the pass's load / store addressing).  The logical state does not change; an X frame on the rank bits may
        remain (it only relabels which rank holds which shard, see _logical_rank_order).  Needed where the reference's
        semantics depend on the index ORDER: the sequential CDF of sample() (reference src/Simulator.cu:164-185)."""
        n, nl = self.n, self.nl
        perm, frame = list(self.perm), self.frame

        def swap(g, l):
            nonlocal frame
            self.engine.swap(g, l)
            inv = {perm[q]: q for q in range(n)}
            qg, ql = inv[g], inv[l]
            perm[qg], perm[ql] = l, g
            bg, bl = (frame >> g) & 1, (frame >> l) & 1
            frame = (frame & ~((1 << g) | (1 << l))) | (bl << g) | (bg << l)

        for g in range(nl, n):
            if perm[g] == g:
                continue
            if perm[g] >= nl:              # sits in another rank bit: bring it down to a local position first
                swap(perm[g], nl - 1)
            swap(g, perm[g])
        # local part: sort the positions with SWAP gates on PHYSICAL qubits (cycle sort), plus any X frame that the
        # exchanges moved onto local bits — one program, applied by its addressing
        inv = {perm[q]: q for q in range(n)}
        recs = []
        for pos in range(nl):
            while inv[pos] != pos:
                q = inv[pos]                                   # belongs at position q
                recs.append((15, pos, q, -1, 0.0))             # SWAP(pos, q)
                inv[pos], inv[q] = inv[q], q
        if recs or (frame & ((1 << nl) - 1)):
            h, info = self.engine.compile_gates(np.array(recs, dtype=GATE_DTYPE) if recs else np.zeros(0, GATE_DTYPE), frame)
            try:
                self.engine.run_program(h)
            finally:
                self.engine.free_program(h)
            frame = info["global_xor"] << nl
        self.perm = list(range(n))
        self.frame = frame
        self._order_preserving = True
        self._pristine = False

    def measure_bit(self, bit: int, uniform: float):
        """Measures LOGICAL index bit `bit` with the caller's uniform draw r: outcome 0 iff r < P(bit = 0)
        (reference src/StateVector.cu:284-313), then collapses and renormalises every shard.  Returns (outcome, p0).
        P0 is the sum of the shards' partial sums in frame-resolved rank order."""
        if not 0 <= bit < self.n:
            raise _lib.InvalidArgument("Qubit index out of range")
        pos = self.perm[bit]
        fx = self.frame >> self.nl
        if pos < self.nl:
            part = self.engine.partial_probability(pos)
            if (self.frame >> pos) & 1:        # pending X on that local bit (only between the steps of a plan; kept for safety)
                part = self.engine.partial_probability(-1) - part
        else:
            mine = ((self.rank ^ fx) >> (pos - self.nl)) & 1
            part = self.engine.partial_probability(-1) if mine == 0 else 0.0
        parts = self._allgather(part)                          # indexed by physical rank
        p0 = 0.0
        for r in range(self.world):                            # frame-resolved order, fixed on every rank
            p0 += parts[r ^ fx]
        outcome = 0 if uniform < p0 else 1
        prob = p0 if outcome == 0 else 1.0 - p0
        if prob <= 0.0:
            raise _lib.QsimError("Measurement outcome has zero probability")
        scale = 1.0 / np.sqrt(prob)
        if pos < self.nl:
            self.engine.collapse(pos, outcome ^ ((self.frame >> pos) & 1), scale)
        else:
            mine = ((self.rank ^ fx) >> (pos - self.nl)) & 1
            self.engine.collapse(-1, 0, scale if mine == outcome else 0.0)
        self._pristine = False
        return outcome, p0

    def measure_qubit(self, qubit: int, uniform: Optional[float] = None) -> int:
        """Simulator::measureQubit: measures index bit n-1-qubit (the reference's big-endian quirk, src/StateVector.cu:87-89)."""
        if not 0 <= qubit < self.n:
            raise _lib.InvalidArgument("Qubit index out of range")
        r = float(np.random.random()) if uniform is None else float(uniform)
        if self.world > 1 and uniform is None:                 # every rank must use the same draw
            r = self._allgather(r)[0]
        return self.measure_bit(self.n - 1 - qubit, r)[0]

    def sample(self, n_shots: int = 0, uniforms: Optional[np.ndarray] = None, seed: Optional[int] = None) -> np.ndarray:
        """Bit-exact distributed sampling in the reference's LOGICAL index order (src/Simulator.cu:164-185: inclusive
        sequential prefix sum over the basis states, lower_bound per shot): if exchanges have permuted the stored
        layout, the identity layout is restored first (restore_identity_layout), then the sequential CDF is chained
        through the shards in rank order.  Returns logical basis-state indices, identical on every rank."""
        if uniforms is None:
            rs = np.random.RandomState(seed)
            uniforms = rs.random_sample(n_shots)
        u = np.ascontiguousarray(uniforms, np.float64)
        if not self._order_preserving and self.perm != list(range(self.n)):
            self.restore_identity_layout()
        fx = self.frame >> self.nl
        my_pos = self.rank ^ fx                       # position of my shard in the frame-resolved order
        if self.world > 1 and hasattr(self.engine, "cdf_prepare"):
            # every shard sweeps its amplitudes now, at the same time; the chain below only stitches and samples
            totals = self._allgather(self.engine.cdf_prepare())           # indexed by physical rank
            self.engine.cdf_classify(float(sum(totals[p ^ fx] for p in range(my_pos))))
        c = 0.0
        result = np.full(len(u), -1, np.int64)
        for pos in range(self.world):                 # chain: shard `pos` continues from the exact sum so far
            if pos == my_pos:
                local, c_end = self.engine.shard_sample(c, pos == 0, u)
                hit = local >= 0
                result[hit] = (np.int64(pos) << np.int64(self.nl)) | local[hit]
            else:
                c_end = 0.0
            ends = self._allgather(c_end if pos == my_pos else -1.0)
            c = max(ends)
        result = self._allreduce_max(result)
        missing = result < 0
        out = self.physical_to_logical_index(np.where(missing, 0, result).astype(np.uint64)).astype(np.int64)
        out[missing] = 1 << self.n                    # past the end, as the reference's lower_bound
        return out

    def get_state_vector(self) -> np.ndarray:
        """Full logical state on every rank (small states only: tests)."""
        import torch
        local = self.engine.local_state()
        if self.world > 1:
            if hasattr(self.engine, "allgather_array"):
                shards = self.engine.allgather_array(local)
            else:
                import torch.distributed as dist
                t = torch.from_numpy(local).cuda()
                outs = [torch.empty_like(t) for _ in range(self.world)]
                dist.all_gather(outs, t)
                shards = [o.cpu().numpy() for o in outs]
        else:
            shards = [local]
        fx = self.frame >> self.nl
        stored = np.concatenate([shards[p ^ fx] for p in range(self.world)])   # frame-resolved physical order
        idx = self.physical_to_logical_index(np.arange(1 << self.n, dtype=np.uint64))
        out = np.empty(1 << self.n, np.complex128)
        out[idx.astype(np.int64)] = stored
        return out

    def close(self):
        self.engine.close()


# ---------------------------------------------------------------------------------------------------------------------
# The C++ driver (qsim::ShardedSimulator, include/qsim/sharded_simulator.hpp) behind the same Python surface
# ---------------------------------------------------------------------------------------------------------------------

def plan_circuit_native(num_qubits: int, n_global: int, gates: np.ndarray, perm: Optional[Sequence[int]] = None,
                        choose_layout: bool = False):
    """The C++ planner (csrc/sharded_plan.cpp) through the C ABI: returns (Plan, start permutation).  Host logic only."""
    g = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
    cap = 2 * len(g) + 2
    steps = np.zeros(3 * cap, np.int64)
    gout = np.zeros(max(len(g), 1), GATE_DTYPE)
    p_in = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    p_start, p_end = np.zeros(num_qubits, np.int32), np.zeros(num_qubits, np.int32)
    n_steps = c_int64()
    _lib.check(_lib.lib().qsim_sharded_plan_circuit(
        int(num_qubits), int(n_global), _lib.gates_ptr(g) if len(g) else None, len(g),
        None if p_in is None else p_in.ctypes.data_as(c_void_p), int(choose_layout), steps.ctypes.data_as(c_void_p), cap,
        gout.ctypes.data_as(c_void_p), p_start.ctypes.data_as(c_void_p), p_end.ctypes.data_as(c_void_p), byref(n_steps)))
    plan = Plan(num_qubits, n_global, n_gates=len(g))
    gi = 0
    for i in range(n_steps.value):
        kind, a, b = (int(x) for x in steps[3 * i:3 * i + 3])
        if kind == 0:
            plan.steps.append(Step("gates", gates=gout[gi:gi + a].copy()))
            gi += a
        else:
            plan.steps.append(Step("swap", global_qubit=a, local_qubit=b))
    plan.perm = [int(x) for x in p_end]
    return plan, [int(x) for x in p_start]


class _NativePlan:
    def __init__(self, handle):
        self._h = handle
        info = (c_int64 * 8)()
        _lib.check(_lib.lib().qsim_sharded_plan_info(handle, info))
        self.n_passes, self.n_ops, self.n_swaps = int(info[0]), int(info[1]), int(info[2])

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:   # (_lib is None while the interpreter shuts down)
            _lib.lib().qsim_sharded_plan_destroy(self._h)
            self._h = c_void_p()


class NativeShardedSimulator:
    """Python mirror of qsim::ShardedSimulator: planning, layout bookkeeping, CUDA-IPC / NCCL exchanges and the distributed
    read-out all run in C++ (NCCL called directly); torch.distributed is used ONCE, to hand rank 0's NCCL unique id to the
    other ranks.  Same surface as ShardedSimulator above (the pure-Python driver, kept for the gloo tests)."""

    def __init__(self, num_qubits: int, exchange: str = "auto"):
        import torch
        import torch.distributed as dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.n = int(num_qubits)
        ident = (ctypes.c_ubyte * 128)()
        if self.world > 1:
            if self.rank == 0:
                _lib.check(_lib.lib().qsim_sharded_unique_id(ident))
            box = [bytes(ident)]
            dist.broadcast_object_list(box, src=0)
            ident = (ctypes.c_ubyte * 128).from_buffer_copy(box[0])
        self._h = c_void_p()
        mode = {"auto": 0, "p2p": 1, "nccl": 2}[exchange]
        _lib.check(_lib.lib().qsim_sharded_create(self.n, self.rank, self.world, ident, mode, byref(self._h)))
        self.stream = torch.cuda.current_stream()
        _lib.check(_lib.lib().qsim_sharded_set_stream(self._h, c_void_p(self.stream.cuda_stream)))
        info = self._info()
        self.nl, self.ng = int(info[0]), int(info[1])
        from .simulator import Simulator

        class _Borrowed(Simulator):          # the shard as a Simulator (timing, counters); owned by the C++ object
            def __del__(self): pass
            close = __del__
        self.local = _Borrowed(self.nl, _handle=c_void_p(_lib.lib().qsim_sharded_local(self._h)))

    def _info(self):
        info = (c_int64 * 8)()
        _lib.check(_lib.lib().qsim_sharded_info(self._h, info))
        return info

    @property
    def fused_exchanges(self) -> int: return int(self._info()[2])
    @property
    def separate_exchanges(self) -> int: return int(self._info()[3])
    @property
    def inplace_exchanges(self) -> int:
        c = c_int64()
        _lib.check(_lib.lib().qsim_sharded_inplace_exchanges(self._h, byref(c)))
        return int(c.value)
    @property
    def split_exchanges(self) -> int:
        c = c_int64()
        _lib.check(_lib.lib().qsim_sharded_split_exchanges(self._h, byref(c)))
        return int(c.value)
    @property
    def exchange(self) -> str: return {0: "none", 1: "p2p", 2: "nccl"}[int(self._info()[4])]
    @property
    def perm(self) -> List[int]:
        out = np.zeros(self.n, np.int32)
        _lib.check(_lib.lib().qsim_sharded_layout(self._h, out.ctypes.data_as(c_void_p), None))
        return [int(x) for x in out]
    @property
    def frame(self) -> int:
        f = c_uint64()
        _lib.check(_lib.lib().qsim_sharded_layout(self._h, None, byref(f)))
        return int(f.value)

    def reset(self): _lib.check(_lib.lib().qsim_sharded_reset(self._h))
    def identity_layout_only(self, on: bool = True): _lib.check(_lib.lib().qsim_sharded_set_identity_layout_only(self._h, int(on)))
    def synchronize(self): _lib.check(_lib.lib().qsim_sharded_synchronize(self._h))
    def barrier(self): _lib.check(_lib.lib().qsim_sharded_barrier(self._h))
    def swap(self, global_position: int, local_position: int): _lib.check(_lib.lib().qsim_sharded_swap(self._h, global_position, local_position))
    def restore_identity_layout(self): _lib.check(_lib.lib().qsim_sharded_restore_identity_layout(self._h))
    def relabel_identity(self): _lib.check(_lib.lib().qsim_sharded_relabel_identity(self._h))
    @property
    def has_second_buffer(self) -> bool: return bool(self._info()[7])

    def run(self, circuit: Circuit):
        g = circuit.gates
        _lib.check(_lib.lib().qsim_sharded_run(self._h, circuit.get_num_qubits(), _lib.gates_ptr(g) if len(g) else None, len(g)))

    def compile_sequence(self, circuit: Circuit, k: int) -> List[_NativePlan]:
        g = circuit.gates
        out = (c_void_p * k)()
        _lib.check(_lib.lib().qsim_sharded_compile(self._h, circuit.get_num_qubits(), _lib.gates_ptr(g) if len(g) else None, len(g),
                                                   int(k), out))
        return [_NativePlan(c_void_p(out[i])) for i in range(k)]

    def compile(self, circuit: Circuit) -> _NativePlan:
        return self.compile_sequence(circuit, 1)[0]

    def execute(self, plan: _NativePlan): _lib.check(_lib.lib().qsim_sharded_execute(self._h, plan._h))
    def release(self, plan): pass

    def sample(self, n_shots: int = 0, uniforms: Optional[np.ndarray] = None, seed: Optional[int] = None) -> np.ndarray:
        if uniforms is None:
            uniforms = np.random.RandomState(seed).random_sample(n_shots)
        u = np.ascontiguousarray(uniforms, np.float64)
        out = np.empty(len(u), np.int64)
        _lib.check(_lib.lib().qsim_sharded_sample(self._h, u.ctypes.data_as(c_void_p), len(u), out.ctypes.data_as(c_void_p)))
        return out

    def measure_qubit(self, qubit: int, uniform: float) -> int:
        res = ctypes.c_int()
        _lib.check(_lib.lib().qsim_sharded_measure(self._h, int(qubit), float(uniform), byref(res)))
        return res.value

    def marginal(self, qubits) -> np.ndarray:
        qs = np.ascontiguousarray(qubits, dtype=np.int32)
        out = np.empty(1 << len(qs), np.float64)
        _lib.check(_lib.lib().qsim_sharded_marginal(self._h, qs.ctypes.data_as(c_void_p), len(qs), out.ctypes.data_as(c_void_p)))
        return out

    def get_total_probability(self) -> float:
        v = c_double()
        _lib.check(_lib.lib().qsim_sharded_total_probability(self._h, byref(v)))
        return v.value

    def local_state(self) -> np.ndarray:
        out = np.empty(1 << self.nl, np.complex128)
        _lib.check(_lib.lib().qsim_sharded_get_local_state(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def get_state_vector(self) -> np.ndarray:
        """Full logical state on every rank (small states only: tests)."""
        import torch
        import torch.distributed as dist
        local = self.local_state()
        if self.world > 1:
            t = torch.from_numpy(local).cuda()
            outs = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(outs, t)
            shards = [o.cpu().numpy() for o in outs]
        else:
            shards = [local]
        perm, fx = self.perm, self.frame >> self.nl
        stored = np.concatenate([shards[p ^ fx] for p in range(self.world)])
        phys = np.arange(1 << self.n, dtype=np.uint64)
        idx = np.zeros_like(phys)
        for q in range(self.n):
            idx |= ((phys >> np.uint64(perm[q])) & np.uint64(1)) << np.uint64(q)
        out = np.empty(1 << self.n, np.complex128)
        out[idx.astype(np.int64)] = stored
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.local._h = c_void_p()
            _lib.lib().qsim_sharded_destroy(self._h)
            self._h = c_void_p()

    __del__ = close
