"""Graph-replay window runner: the same model code as WindowRunner, executed in STATIC-CAPACITY mode so a
whole frame is one CUDA graph launch.

Why: the eager path issues ~250 launches per frame from Python and reads the vertex counters back three
times per frame to size its tensors -- at ~10 ms of GPU work per window the host
is as slow as the device.  Here every per-point / per-vertex tensor has a fixed capacity as its row
count, the live counts stay in device memory (`*_dev` arguments of the C ABI, looked up by capacity in
_lib's registry), nothing synchronises, and each KIND of frame (first / middle / last of a window) is
captured once and replayed:  host work per frame = two input copies + one graph launch.

Hidden states live in persistent buffers (fusion._HiddenState), the lattice objects of all levels are
reused across windows (Lattice.reset empties them), so the captured addresses stay valid.

Safety net: after each window the real vertex counts are compared with the capacities (they ride back
with the labels); a window that outgrew them is re-run on the eager path, and so is any configuration
whose PointNet widths are not the fused front end's.
"""
import torch

from . import _lib, ops
from .lattice import Lattice
from .runner import WindowRunner

_STATIC_FUSION = ("gru", "lstm", "aflow", "linear", "maxpool", "cga", "none")


def _round_up(x, m):
    return (int(x) + m - 1) // m * m


class GraphWindowRunner(WindowRunner):
    def __init__(self, cfg_path, nr_classes=26, device=None, headroom=1.35, operands="f16"):
        super().__init__(cfg_path, nr_classes, device, operands)
        self.headroom = headroom
        self.graphs = {}
        self.kernels = {}      # frame kind -> kernels of this library inside its graph (bench.py: gpu_launches)
        self.caps = None
        # static capacities need every per-row kernel to take its live count from the device: that holds for the FUSED
        # PointNet front end only (widths [16,32,64], one value column); other cfgs take the eager path, where the
        # unfused scatter_max / scatter_add see exactly-sized tensors
        self.supported = (all(k in _STATIC_FUSION for k in self.model.rnn_modules) and self.model.sequence_learning
                          and self.model.point_net_seq.layer_widths == [16, 32, 64])
        self.fallbacks = 0
        self._force_eager = False

    # ---- capacities from an eager probe window ---------------------------------------------------------
    def plan(self, frames_dev):
        """runs one eager window to learn the sizes, then fixes the capacities (with headroom)"""
        if any(v.shape[1] != 1 for _, v in frames_dev):
            self.supported = False     # the fused PointNet front end takes [x, y, z, value] rows only
        out = super().infer_window_device(frames_dev)
        counts, lvl = [], self.lattice
        while lvl is not None:
            counts.append(lvl.nr_lattice_vertices())
            lvl = lvl._coarse
        n_max = max(p.shape[0] for p, _ in frames_dev)
        caps = {"n": _round_up(n_max * 1.02 + 1024, 4096)}
        vcaps = [_round_up(c * self.headroom + 512, 1024) for c in counts]
        vcaps = [min(v, self.lattice.capacity) for v in vcaps]
        # row classes must be told apart by their capacity alone: collisions move up by 1024 while that stays within the
        # hash table's own capacity (the table clamps its counter there, and the overflow check `count < cap` must be
        # able to fail), else down
        used = {caps["n"], 4 * caps["n"]}
        for i, v in enumerate(vcaps):
            up = v
            while up in used:
                up += 1024
            if up <= self.lattice.capacity:
                v = up
            else:
                while v in used and v > 1024:
                    v -= 1024
                if v in used or v <= counts[i]:
                    self.supported = False   # no distinct capacity left below the table's: such windows run eagerly
            used.add(v)
            vcaps[i] = v
        caps["v"] = vcaps
        self.caps = caps
        self._alloc_static()
        return out

    def _alloc_static(self):
        dev, n = self.device, self.caps["n"]
        self.pos_buf = torch.zeros(n, 3, dtype=torch.float32, device=dev)
        self.val_buf = torch.zeros(n, 1, dtype=torch.float32, device=dev)
        self.sizes = torch.zeros(2, dtype=torch.int32, device=dev)          # [n, 4n]
        self.static_lattice = Lattice.create(self.cfg_path, "lattice")
        self.static_lattice.set_static(self.caps["v"])
        # materialise the coarse levels now so their tables exist before any capture
        lvl = self.static_lattice
        for _ in range(len(self.caps["v"]) - 1):
            lvl = lvl.coarse_level()
        self._counts_host = torch.zeros(4 * len(self.caps["v"]) + 1, dtype=torch.int32).pin_memory()   # + the fp16 range flag
        self.graphs = {}
        self.pool = None

    def _registry(self):
        reg = {self.caps["n"]: self.sizes[0:1], 4 * self.caps["n"]: self.sizes[1:2]}
        lvl = self.static_lattice
        for cap in self.caps["v"]:
            reg[cap] = lvl.hash_table.count_tensor()
            lvl = lvl._coarse
        return reg

    # ---- one frame ---------------------------------------------------------------------------------------
    def _frame(self, kind):
        first, last = kind
        if first:
            self.model.reset_sequence()
        with ops.tc_operands(self.operands, self.range_flag):   # the flag's address is baked into the captured graph
            out, _, _ = self.model(self.static_lattice, self.pos_buf, self.val_buf, not last, False)
        return out

    def _run_frame(self, kind, p, v):
        n = p.shape[0]
        if n > self.caps["n"]:
            raise OverflowError("frame has more points than the static capacity")
        self.pos_buf[:n].copy_(p, non_blocking=True)
        self.val_buf[:n].copy_(v, non_blocking=True)
        _lib.check(_lib.load().ltn_set_int2(_lib.ptr(self.sizes), n, 4 * n, _lib.stream()), "ltn_set_int2")
        g = self.graphs.get(kind)
        if g is None:
            _lib.set_static_rows(self._registry())
            try:
                with torch.no_grad():
                    # eager run in static mode first: creates the lazy buffers / weight splits outside the
                    # capture and EXECUTES the frame (capturing records work without running it), so the
                    # recurrent state is the one the next frame of this capture window expects
                    self._frame(kind)
                    torch.cuda.current_stream().synchronize()
                    graph = torch.cuda.CUDAGraph()
                    l0 = _lib.load().ltn_launch_count()
                    with torch.cuda.graph(graph, pool=self.pool):
                        out = self._frame(kind)
                    self.kernels[kind] = _lib.load().ltn_launch_count() - l0
                    if self.pool is None:
                        self.pool = graph.pool()
            finally:
                _lib.set_static_rows(None)
            self.graphs[kind] = (graph, out)
            return None
        g[0].replay()
        return g[1]

    def capture(self, frames_dev):
        """Captures the graphs of every frame kind on a throw-away window.  Each kind is first run eagerly
        in static mode (its outputs are discarded), which advances the recurrent state exactly like the
        real frame would, so the following kinds see a consistent state; the captured graphs themselves
        only record work and do not execute."""
        if self.caps is None:
            self.plan(frames_dev)
        if not self.supported:
            return self
        T = len(frames_dev)
        for t, (p, v) in enumerate(frames_dev):
            self._run_frame((t == 0, t == T - 1), p, v)   # captures the kinds seen for the first time, replays the others
        torch.cuda.synchronize()
        return self

    def infer_window_device(self, frames_dev):
        """NOTE on lifetime: in graph mode the result is a view of the captured graph's static output tensor -- valid
        until the next window runs on this runner; clone it to keep it (infer_window's labels likewise live in a reused
        pinned buffer)."""
        if not self.supported or self.caps is None or self._force_eager:
            return super().infer_window_device(frames_dev)
        if any(p.shape[0] > self.caps["n"] for p, _ in frames_dev):   # more points than the static buffers hold
            self.fallbacks += 1
            return WindowRunner.infer_window_device(self, frames_dev)
        T = len(frames_dev)
        kinds = [(t == 0, t == T - 1) for t in range(T)]
        if any(k not in self.graphs for k in kinds):
            self.capture(frames_dev)
        out = None
        for kind, (p, v) in zip(kinds, frames_dev):
            out = self._run_frame(kind, p, v)
        self._last_n = frames_dev[-1][0].shape[0]
        return out[: self._last_n]

    def kernels_per_window(self, nr_frames):
        return sum(self.kernels.get((t == 0, t == nr_frames - 1), 0) for t in range(nr_frames))

    def _copy_checks(self, host):
        """queues the device->host copies of the window's safety checks: the four counters of every level (vertices,
        previous, overflowed, keys out of the packable range) + the fp16 range flag"""
        lvl, i, nv = self.static_lattice, 0, len(self.caps["v"])
        while lvl is not None and i < nv:
            host[4 * i:4 * i + 4].copy_(lvl.hash_table.counters[0:4], non_blocking=True)
            lvl, i = lvl._coarse, i + 1
        host[4 * nv:4 * nv + 1].copy_(self.range_flag, non_blocking=True)

    def _checks_ok(self, host):
        vals, nv = host.tolist(), len(self.caps["v"])
        if vals[4 * nv]:   # an activation left the fp16 range: clear the flag, the caller re-runs the window
            self.range_flag.zero_()
            self.range_fallbacks += 1
            return False
        for i, cap in enumerate(self.caps["v"]):
            count, _, overflowed, out_of_range = vals[4 * i:4 * i + 4]
            if out_of_range:   # the eager re-run raises the RuntimeError of HashTable._sync_counters for it
                return False
            if overflowed or not int(count) < cap:
                return False
        return True

    def counts_ok(self):
        """one small device->host read: did every level stay within its capacity, and every staged activation
        within the fp16 range?"""
        self._copy_checks(self._counts_host)
        torch.cuda.current_stream().synchronize()
        return self._checks_ok(self._counts_host)

    def infer_window(self, frames_host):
        if not self.supported or self.caps is None:
            return super().infer_window(frames_host)
        dev = self.device
        frames = [(p.to(dev, non_blocking=True), v.to(dev, non_blocking=True)) for p, v in frames_host]
        out = self.infer_window_device(frames)
        labels = out.argmax(1)
        if self._labels_host is None or self._labels_host.shape[0] < labels.shape[0]:
            self._labels_host = torch.empty(labels.shape[0], dtype=torch.int64).pin_memory()
        host = self._labels_host[: labels.shape[0]]
        host.copy_(labels, non_blocking=True)
        if not self.counts_ok():          # synchronises; the labels have landed too
            self.fallbacks += 1
            self._force_eager = True
            try:
                return super().infer_window(frames_host)
            finally:
                self._force_eager = False
        return host


class MultiWindowRunner:
    """`lanes` independent windows in flight on one GPU, each on its own CUDA stream with its own
    GraphWindowRunner (own lattice, hidden states, captured graphs; identical weights).

    Windows are independent by construction (a fresh lattice per window, train_ln.py:236-239), and a single
    window cannot fill a B200: the coarse-level kernels launch 17-60 CTAs on 148 SMs and every frame is a
    strictly sequential chain.  Interleaving the frame graphs of two windows lets one window's small kernels
    run on the SMs the other leaves idle -- the intra-GPU form of the sharding-by-window of SURVEY.md 8(e)."""

    def __init__(self, cfg_path, nr_classes=26, device=None, lanes=4, operands="f16"):
        self.lanes = [GraphWindowRunner(cfg_path, nr_classes, device, operands=operands) for _ in range(lanes)]
        self.device = self.lanes[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(lanes)]
        self.supported = self.lanes[0].supported

    def prepare(self, frames_dev, state_dict_fn, plan_windows=None):
        for lane, s in zip(self.lanes, self.streams):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                lane.materialise_parameters(frames_dev, state_dict_fn)
                for w in (plan_windows or [frames_dev]):
                    if lane.caps is None or max(p.shape[0] for p, _ in w) > lane.caps["n"]:
                        lane.plan(w)
                lane.capture(frames_dev)
            torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.supported = all(l.supported for l in self.lanes)
        return self

    def infer_windows_device(self, windows_dev):
        """windows_dev: up to `lanes` windows (lists of (positions, values) CUDA tensors).  Frame t of every
        window is issued before frame t+1 of any, so the streams interleave.  Returns the log-softmax of each
        window's last frame; the caller's stream waits for all lanes.  The returned tensors are views of each lane's static
        graph output: valid until that lane runs its next window (clone to keep)."""
        cur = torch.cuda.current_stream()
        outs = [None] * len(windows_dev)
        T = max(len(w) for w in windows_dev)
        for s in self.streams[: len(windows_dev)]:
            s.wait_stream(cur)
        for t in range(T):
            for i, w in enumerate(windows_dev):
                if t < len(w):
                    with torch.cuda.stream(self.streams[i]):
                        lane = self.lanes[i]
                        kind = (t == 0, t == len(w) - 1)
                        if kind not in lane.graphs:
                            lane.capture(w)
                        outs[i] = lane._run_frame(kind, w[t][0], w[t][1])
        for i, w in enumerate(windows_dev):
            outs[i] = outs[i][: w[-1][0].shape[0]]
            cur.wait_stream(self.streams[i])
        return outs

    # ---- end to end: pinned host buffers in, labels out; submission and collection are decoupled so the next
    # group of windows is already queued while the host waits for the previous one -----------------------------
    def submit(self, windows_host):
        """queues up to `lanes` windows (pinned host tensors); returns a ticket for collect()"""
        dev = self.device
        cur = torch.cuda.current_stream()
        slot = getattr(self, "_slot", 0)
        self._slot = slot ^ 1
        if not hasattr(self, "_host_out"):
            self._host_out = {}
        wins, ticket = [], {"host": windows_host, "labels": [], "counts": [], "events": []}
        oversize = [any(p.shape[0] > self.lanes[i].caps["n"] for p, _ in w) for i, w in enumerate(windows_host)]
        for i, w in enumerate(windows_host):   # a frame kind this lane has not captured yet (e.g. a 1-frame window)
            kinds = {(t == 0, t == len(w) - 1) for t in range(len(w))}
            if not oversize[i] and any(k not in self.lanes[i].graphs for k in kinds):
                with torch.cuda.stream(self.streams[i]):
                    self.lanes[i].capture([(p.to(dev), v.to(dev)) for p, v in w])
        for i, w in enumerate(windows_host):
            self.streams[i].wait_stream(cur)
            with torch.cuda.stream(self.streams[i]):
                wins.append([] if oversize[i] else [(p.to(dev, non_blocking=True), v.to(dev, non_blocking=True)) for p, v in w])
        T = max(len(w) for w in windows_host)
        outs = [None] * len(wins)
        for t in range(T):
            for i, w in enumerate(wins):
                if t < len(w):
                    with torch.cuda.stream(self.streams[i]):
                        outs[i] = self.lanes[i]._run_frame((t == 0, t == len(w) - 1), w[t][0], w[t][1])
        for i, w in enumerate(wins):
            if oversize[i]:   # collect() runs it on the eager path
                ticket["labels"].append(None)
                ticket["counts"].append(None)
                ticket["events"].append(None)
                continue
            lane, n = self.lanes[i], w[-1][0].shape[0]
            with torch.cuda.stream(self.streams[i]):
                lab = outs[i][:n].argmax(1)
                key = (i, slot)
                buf = self._host_out.get(key)
                if buf is None or buf[0].shape[0] < n:
                    buf = (torch.empty(max(n, lane.caps["n"]), dtype=torch.int64).pin_memory(),
                           torch.zeros(4 * len(lane.caps["v"]) + 1, dtype=torch.int32).pin_memory())
                    self._host_out[key] = buf
                buf[0][:n].copy_(lab, non_blocking=True)
                lane._copy_checks(buf[1])
                ev = torch.cuda.Event()
                ev.record()
            ticket["labels"].append(buf[0][:n])
            ticket["counts"].append(buf[1])
            ticket["events"].append(ev)
        return ticket

    def collect(self, ticket):
        """waits for a ticket's windows; returns their predicted labels (pinned host int64 tensors, valid until
        the slot is reused two submissions later).  A window that outgrew the static capacities is re-run on
        the eager path."""
        out = []
        for i, (lab, counts, ev) in enumerate(zip(ticket["labels"], ticket["counts"], ticket["events"])):
            lane = self.lanes[i]
            ok = False
            if ev is not None:
                ev.synchronize()
                with torch.cuda.stream(self.streams[i]):
                    ok = lane._checks_ok(counts)
            if ok:
                out.append(lab)
            else:
                lane.fallbacks += 1
                lane._force_eager = True
                try:
                    with torch.cuda.stream(self.streams[i]):
                        out.append(WindowRunner.infer_window(lane, ticket["host"][i]).clone())
                finally:
                    lane._force_eager = False
        return out

    def infer_windows(self, windows_host):
        """pinned host buffers in, predicted labels (pinned host int64) out, synchronised"""
        return self.collect(self.submit(windows_host))

    def kernels_per_window(self, nr_frames):
        return self.lanes[0].kernels_per_window(nr_frames)

    def counts_ok(self):
        return all(l.counts_ok() for l in self.lanes)


class _LockstepCoordinator:
    """Runs the frame code of several lanes as host threads that hand a baton around: a lane runs until its next
    batchable convolution (ops.conv_tc -> request()) or the end of its frame, then the next lane runs.  When every lane
    has deposited its request the coordinator (the calling thread) joins the lane streams into the main stream, issues
    ONE batched launch there and lets the lane streams continue behind it.  Only used while a frame is being CAPTURED
    (or warmed up): replays involve no Python at all."""

    def __init__(self, main_stream, lane_streams):
        import threading
        self.main, self.streams = main_stream, lane_streams
        self.n = len(lane_streams)
        self.go = [threading.Semaphore(0) for _ in range(self.n)]
        self.back = threading.Semaphore(0)
        self.pending = [None] * self.n
        self.finished = [False] * self.n
        self.errors = [None] * self.n
        self.results = [None] * self.n
        self.batched_launches = 0
        self.log = []            # (C, S, F) of every batched launch, in issue order
        self._tl = threading.local()

    # ---- lane side -------------------------------------------------------------------------------------------------
    def request(self, req):
        i = self._tl.lane
        ev = torch.cuda.Event()
        ev.record(self.streams[i])          # everything this request reads has been queued on the lane's stream
        self.pending[i] = (req, ev)
        self.back.release()
        self.go[i].acquire()
        done = self.pending[i]
        self.pending[i] = None
        self.streams[i].wait_event(done)    # the lane continues behind the batched launch

    def _lane_main(self, i, fn):
        self._tl.lane = i
        self.go[i].acquire()
        try:
            ops._BATCH.ctx = self
            with torch.cuda.stream(self.streams[i]):
                self.results[i] = fn(i)
        except BaseException as e:  # noqa: BLE001 -- handed to the coordinating thread
            self.errors[i] = e
        finally:
            ops._BATCH.ctx = None
            self.finished[i] = True
            self.back.release()

    # ---- coordinator side ------------------------------------------------------------------------------------------
    def run(self, fn):
        """fn(lane_index) runs in lane i's thread with lane i's stream current; returns the list of results"""
        import threading
        threads = [threading.Thread(target=self._lane_main, args=(i, fn), daemon=True) for i in range(self.n)]
        for t in threads:
            t.start()
        fork = torch.cuda.Event()
        fork.record(self.main)
        for s in self.streams:
            s.wait_event(fork)
        while True:
            for i in range(self.n):
                if not self.finished[i]:
                    self.go[i].release()
                    self.back.acquire()
            if any(e is not None for e in self.errors):
                break
            waiting = [i for i in range(self.n) if not self.finished[i]]
            if not waiting:
                break
            if len(waiting) != self.n:
                self.errors[waiting[0]] = RuntimeError("lock-step lanes diverged: some finished their frame while others wait at a layer")
                break
            for i in waiting:
                self.main.wait_event(self.pending[i][1])
            with torch.cuda.stream(self.main):
                ops.conv_tc_batched([self.pending[i][0] for i in waiting])
                done = torch.cuda.Event()
                done.record(self.main)
            self.batched_launches += 1
            r0 = self.pending[waiting[0]][0]
            self.log.append((int(r0["C"]), int(r0["S"]) if r0["nbr"] is not None else 1, int(r0["F"])))
            for i in waiting:
                self.pending[i] = done
        failed = next((e for e in self.errors if e is not None), None)
        if failed is not None:
            for i in range(self.n):          # let the surviving lane threads run to their end without batching
                while not self.finished[i]:
                    self.pending[i] = torch.cuda.Event()
                    self.go[i].release()
                    self.back.acquire()
            raise failed
        for t in threads:
            t.join()
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            self.main.wait_event(ev)
        return self.results


class LockstepRunner:
    """`lanes` independent windows advance through their frames TOGETHER on one GPU: one CUDA graph per frame kind
    covers all lanes, and every tensor-core layer is ONE persistent launch over the tiles of all lanes
    (csrc/ltn_conv_batched.cu) instead of one partial-wave launch per lane -- SURVEY.md 8(b) "B lattices per launch".
    Everything else of a lane (hash build, PointNet, GroupNorm statistics, gates, AFlow, slicing) stays on that lane's
    own stream inside the graph, so those small kernels of different windows still overlap.

    Each lane is a GraphWindowRunner (own model copy, lattice, hidden-state buffers, static capacities); windows that
    do not fit the static capacities, and groups smaller than `lanes`, go through the lanes' own per-window paths."""

    TRACE_RECORDS = 192      # batched launches per frame kind that can be traced
    TRACE_STRIDE = 2 + 2 * 148

    def __init__(self, cfg_path, nr_classes=26, device=None, lanes=4, operands="f16", trace=True):
        if lanes < 1 or lanes > ops.MAX_BATCH:
            raise RuntimeError("between 1 and %d lanes" % ops.MAX_BATCH)
        self.trace = trace       # every batched launch stamps its CTAs' entry / exit times into a per-kind record (trace_group)
        self.trace_bufs = {}
        self.layer_log = {}
        self.lanes = [GraphWindowRunner(cfg_path, nr_classes, device, operands=operands) for _ in range(lanes)]
        self.device = self.lanes[0].device
        self.main = torch.cuda.Stream(device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(lanes)]
        self.supported = self.lanes[0].supported
        self.graphs = {}       # frame kind -> (graph, [per-lane static output])
        self.kernels = {}      # frame kind -> kernel launches inside the graph (all lanes)
        self.batched = {}      # frame kind -> batched tensor-core launches among them
        self.pool = None
        self._host_out = {}
        self._slot = 0

    # ---- preparation -----------------------------------------------------------------------------------------------
    def prepare(self, frames_dev, state_dict_fn, plan_windows=None):
        for lane in self.lanes:
            lane.materialise_parameters(frames_dev, state_dict_fn)
            for w in (plan_windows or [frames_dev]):
                if lane.caps is None or max(p.shape[0] for p, _ in w) > lane.caps["n"]:
                    lane.plan(w)
        self.supported = all(l.supported for l in self.lanes)
        torch.cuda.synchronize()
        if self.supported:
            self.capture([frames_dev] * len(self.lanes))
        return self

    def _load_inputs(self, lane, p, v):
        n = p.shape[0]
        if n > lane.caps["n"]:
            raise OverflowError("frame has more points than the static capacity")
        lane.pos_buf[:n].copy_(p, non_blocking=True)
        lane.val_buf[:n].copy_(v, non_blocking=True)
        _lib.check(_lib.load().ltn_set_int2(_lib.ptr(lane.sizes), n, 4 * n, _lib.stream()), "ltn_set_int2")

    def _lane_frame(self, i, kind):
        lane = self.lanes[i]
        _lib.set_static_rows(lane._registry())
        try:
            with torch.no_grad():
                return lane._frame(kind)
        finally:
            _lib.set_static_rows(None)

    def _capture_kind(self, kind):
        """the lanes' inputs are loaded; warms the kind up (eager, static mode, lock-step) and captures it"""
        cur = torch.cuda.current_stream()
        self.main.wait_stream(cur)
        with torch.cuda.stream(self.main):
            # executed once outside the capture: creates lazy buffers / weight splits and ADVANCES the recurrent state
            # like the real frame would, so the kinds captured next see a consistent state
            _LockstepCoordinator(self.main, self.streams).run(lambda i: self._lane_frame(i, kind))
            self.main.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = _lib.load().ltn_launch_count()
            coord = _LockstepCoordinator(self.main, self.streams)
            if self.trace:
                self.trace_bufs[kind] = torch.zeros(self.TRACE_RECORDS, self.TRACE_STRIDE, dtype=torch.int64, device=self.device)
                _lib.load().ltn_conv_batched_trace(_lib.ptr(self.trace_bufs[kind]), self.TRACE_RECORDS)
            try:
                with torch.cuda.graph(graph, pool=self.pool, stream=self.main, capture_error_mode="thread_local"):
                    outs = coord.run(lambda i: self._lane_frame(i, kind))
            finally:
                if self.trace:
                    _lib.load().ltn_conv_batched_trace(None, 0)
            self.layer_log[kind] = list(coord.log)
            self.kernels[kind] = _lib.load().ltn_launch_count() - l0
            self.batched[kind] = coord.batched_launches
            if self.pool is None:
                self.pool = graph.pool()
        cur.wait_stream(self.main)
        self.graphs[kind] = (graph, outs)

    def capture(self, windows_dev):
        """captures every frame kind of these windows (one per lane, equal frame counts) that is not captured yet"""
        T = len(windows_dev[0])
        for t in range(T):
            kind = (t == 0, t == T - 1)
            with torch.cuda.stream(self.main):
                for lane, w in zip(self.lanes, windows_dev):
                    self._load_inputs(lane, w[t][0], w[t][1])
            if kind in self.graphs:
                with torch.cuda.stream(self.main):
                    self.graphs[kind][0].replay()
            else:
                torch.cuda.current_stream().wait_stream(self.main)
                self._capture_kind(kind)
                # the capture only RECORDED the frame; the warm-up run before it advanced the state once, which is what
                # the following kinds need
        torch.cuda.synchronize()
        return self

    # ---- execution -------------------------------------------------------------------------------------------------
    def _fits(self, windows):
        T = len(windows[0])
        return (self.supported and len(windows) == len(self.lanes) and all(len(w) == T for w in windows)
                and all(p.shape[0] <= lane.caps["n"] for lane, w in zip(self.lanes, windows) for p, _ in w))

    def _run_group(self, windows_dev):
        """queues one full group on the main stream; returns the per-lane static outputs of the last frame"""
        T = len(windows_dev[0])
        kinds = [(t == 0, t == T - 1) for t in range(T)]
        if any(k not in self.graphs for k in kinds):
            self.capture(windows_dev)
        outs = None
        with torch.cuda.stream(self.main):
            for t, kind in enumerate(kinds):
                for lane, w in zip(self.lanes, windows_dev):
                    self._load_inputs(lane, w[t][0], w[t][1])
                graph, outs = self.graphs[kind]
                graph.replay()
        return outs

    def begin_windows_device(self, windows_dev):
        """queues one group behind the caller's current stream WITHOUT making that stream wait for it; returns a handle for
        end_windows_device().  Several runners (GroupedLockstepRunner) begin their groups first and end them afterwards, so
        the groups overlap on the device."""
        cur = torch.cuda.current_stream()
        if not self._fits(windows_dev):
            return ("lanes", [self.lanes[i % len(self.lanes)].infer_window_device(w) for i, w in enumerate(windows_dev)])
        self.main.wait_stream(cur)
        outs = self._run_group(windows_dev)
        return ("group", [o[: w[-1][0].shape[0]] for o, w in zip(outs, windows_dev)])

    def end_windows_device(self, handle):
        if handle[0] == "group":
            torch.cuda.current_stream().wait_stream(self.main)
        return handle[1]

    def infer_windows_device(self, windows_dev):
        """windows_dev: lists of (positions, values) CUDA tensors, one per lane.  Returns the log-softmax of each window's
        last frame (views of the graphs' static outputs: valid until the next group runs; clone to keep).  The caller's
        stream waits for the group."""
        return self.end_windows_device(self.begin_windows_device(windows_dev))

    def submit(self, windows_host):
        """queues one group of windows given as PINNED host tensors; returns a ticket for collect().  The host -> device
        copies run on a copy stream, so the next group's inputs travel while the current group computes."""
        dev = self.device
        ticket = {"host": windows_host, "labels": None, "counts": None, "event": None, "staged": None}
        if not self._fits(windows_host):
            return ticket
        with torch.cuda.stream(self.copy_stream):
            staged = [[(p.to(dev, non_blocking=True), v.to(dev, non_blocking=True)) for p, v in w] for w in windows_host]
            copied = torch.cuda.Event()
            copied.record()
        self.main.wait_event(copied)
        outs = self._run_group(staged)
        slot, self._slot = self._slot, self._slot ^ 1
        labels, counts = [], []
        with torch.cuda.stream(self.main):
            for i, (lane, w) in enumerate(zip(self.lanes, windows_host)):
                n = w[-1][0].shape[0]
                lab = outs[i][:n].argmax(1)
                buf = self._host_out.get((i, slot))
                if buf is None:
                    buf = (torch.empty(lane.caps["n"], dtype=torch.int64).pin_memory(),
                           torch.zeros(4 * len(lane.caps["v"]) + 1, dtype=torch.int32).pin_memory())
                    self._host_out[(i, slot)] = buf
                buf[0][:n].copy_(lab, non_blocking=True)
                lane._copy_checks(buf[1])
                labels.append(buf[0][:n])
                counts.append(buf[1])
            ev = torch.cuda.Event()
            ev.record()
        ticket.update(labels=labels, counts=counts, event=ev, staged=staged)
        return ticket

    def collect(self, ticket):
        """waits for a ticket; returns the predicted labels per window (pinned host int64, valid until the slot is reused
        two submissions later).  Windows that outgrew the static capacities (or raised the fp16 range flag) are re-run on
        the lane's eager path."""
        if ticket["event"] is None:
            return [WindowRunner.infer_window(self.lanes[i % len(self.lanes)], w).clone() for i, w in enumerate(ticket["host"])]
        ticket["event"].synchronize()
        out = []
        for i, lane in enumerate(self.lanes):
            if lane._checks_ok(ticket["counts"][i]):
                out.append(ticket["labels"][i])
            else:
                lane.fallbacks += 1
                lane._force_eager = True
                try:
                    out.append(WindowRunner.infer_window(lane, ticket["host"][i]).clone())
                finally:
                    lane._force_eager = False
        ticket["staged"] = None
        return out

    def infer_windows(self, windows_host):
        return self.collect(self.submit(windows_host))

    def trace_group(self, windows_dev):
        """Runs one full group and returns one record per batched tensor-core launch, measured INSIDE the replayed graphs:
        {kind, C, S, F, rows (live, summed over the windows), tiles, ctas, us (first CTA entry -> last CTA exit), flop}."""
        if not self.trace or not self._fits(windows_dev):
            return []
        T = len(windows_dev[0])
        kinds = [(t == 0, t == T - 1) for t in range(T)]
        out = []
        cur = torch.cuda.current_stream()
        self.main.wait_stream(cur)
        with torch.cuda.stream(self.main):
            for t, kind in enumerate(kinds):
                if kind not in self.graphs:
                    return []
                for lane, w in zip(self.lanes, windows_dev):
                    self._load_inputs(lane, w[t][0], w[t][1])
                self.trace_bufs[kind].zero_()
                self.graphs[kind][0].replay()
                self.main.synchronize()   # the records of this kind are read before the next frame of the same kind overwrites them
                rec = self.trace_bufs[kind].cpu().numpy()
                for j, (C, S, F) in enumerate(self.layer_log[kind][: self.TRACE_RECORDS]):
                    rows, tiles = int(rec[j, 0]), int(rec[j, 1])
                    stamps = rec[j, 2:].reshape(-1, 2)
                    live = stamps[:, 0] > 0
                    if tiles <= 0 or not live.any():
                        continue
                    us = float(stamps[live, 1].max() - stamps[live, 0].min()) / 1e3
                    out.append({"frame": t, "C": C, "S": S, "F": F, "rows": rows, "tiles": tiles, "ctas": int(live.sum()), "us": us,
                                "flop": 2.0 * rows * S * C * F,
                                "cta_us_mean": float((stamps[live, 1] - stamps[live, 0]).mean()) / 1e3,
                                "t0_ns": int(stamps[live, 0].min()), "t1_ns": int(stamps[live, 1].max())})
        cur.wait_stream(self.main)
        return out

    def counts_ok(self):
        return all(l.counts_ok() for l in self.lanes)

    def kernels_per_group(self, nr_frames):
        return sum(self.kernels.get((t == 0, t == nr_frames - 1), 0) for t in range(nr_frames))

    def batched_per_group(self, nr_frames):
        return sum(self.batched.get((t == 0, t == nr_frames - 1), 0) for t in range(nr_frames))


class GroupedLockstepRunner:
    """`groups` lock-step groups of `lanes` windows each, every group with its own graphs and streams, in flight
    together.  A persistent batched launch ends with a tail (CTAs that got one tile fewer, the last epilogue) and the
    next layer's launch cannot start before it: with a second group queued behind, that group's CTAs take the SMs the
    first one releases, so the per-layer fixed costs of one group hide under the other's work.  Same interface as
    LockstepRunner; a full set is groups x lanes windows."""

    def __init__(self, cfg_path, nr_classes=26, device=None, lanes=4, groups=2, operands="f16", trace=True):
        self.groups = [LockstepRunner(cfg_path, nr_classes, device, lanes=lanes, operands=operands, trace=trace) for _ in range(groups)]
        self.device = self.groups[0].device
        self.per_group = lanes
        self.lanes = [l for g in self.groups for l in g.lanes]
        self.supported = all(g.supported for g in self.groups)

    def prepare(self, frames_dev, state_dict_fn, plan_windows=None):
        for g in self.groups:
            g.prepare(frames_dev, state_dict_fn, plan_windows)
        self.supported = all(g.supported for g in self.groups)
        return self

    def _split(self, windows):
        k = self.per_group
        return [windows[i:i + k] for i in range(0, len(windows), k)]

    def infer_windows_device(self, windows_dev):
        handles = [g.begin_windows_device(part) for g, part in zip(self.groups, self._split(windows_dev))]   # all queued first:
        outs = []                                                                                             # the groups overlap
        for g, h in zip(self.groups, handles):
            outs += g.end_windows_device(h)
        return outs

    def submit(self, windows_host):
        return [g.submit(part) for g, part in zip(self.groups, self._split(windows_host))]

    def collect(self, tickets):
        out = []
        for g, t in zip(self.groups, tickets):
            out += g.collect(t)
        return out

    def infer_windows(self, windows_host):
        return self.collect(self.submit(windows_host))

    def counts_ok(self):
        return all(g.counts_ok() for g in self.groups)

    def kernels_per_group(self, nr_frames):
        return sum(g.kernels_per_group(nr_frames) for g in self.groups)

    def trace_group(self, windows_dev):
        """records of ONE group's launches with the other groups running beside it (the timed configuration)"""
        parts = self._split(windows_dev)
        handles = [g.begin_windows_device(part) for g, part in zip(self.groups[1:], parts[1:])]
        rec = self.groups[0].trace_group(parts[0])
        for g, h in zip(self.groups[1:], handles):
            g.end_windows_device(h)
        return rec
