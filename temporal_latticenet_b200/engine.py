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
whose fusion modules have no static-capacity implementation.
"""
import torch

from . import _lib, ops
from .lattice import Lattice
from .runner import WindowRunner

_STATIC_FUSION = ("gru", "lstm", "aflow", "none")


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
