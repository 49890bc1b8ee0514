"""CPU dry run of the segmented-graph data-parallel overlap (Trainer._capture_segments / _replay_segments,
CAPDEC_DP_OVERLAP=2): torch.cuda's graph / stream / event objects and the engine are replaced by recording fakes, so this
checks the HOST ordering logic only — one capture per backward block boundary, the global token count all-reduced first,
the bucket of block l reduced AND its parameters updated right after the segment that finished block l,
[tail | mapper | wte | wpe] last, no collective inside any capture, every collective on the side stream after an event on
the main stream.  The numerical check runs on 2 GPUs (tests/test_dp_gpu.py, which passes on hardware)."""
import contextlib

import torch

LOG = []


class FakeGraph:
    n = 0

    def __init__(self):
        FakeGraph.n += 1
        self.id = FakeGraph.n
        self.open = False

    def capture_begin(self, pool=None):
        assert not any(g.open for g in FakeGraph.live), "nested capture"
        self.open = True
        FakeGraph.live.append(self)
        LOG.append(("begin", self.id))

    def capture_end(self):
        assert self.open
        self.open = False
        LOG.append(("end", self.id))

    def replay(self):
        LOG.append(("replay", self.id))


FakeGraph.live = []


class FakeStream:
    def __init__(self, name="side", device=None):
        self.name = name

    def wait_stream(self, other):
        LOG.append(("wait_stream", self.name, other.name))

    def wait_event(self, ev):
        LOG.append(("wait_event", self.name, ev.on))


class FakeEvent:
    def __init__(self, *a, **k):
        self.on = None

    def record(self, stream=None):
        self.on = (stream or MAIN).name


MAIN = FakeStream("main")
CUR = [MAIN]


@contextlib.contextmanager
def fake_stream_ctx(s):
    CUR.append(s)
    try:
        yield
    finally:
        CUR.pop()


class FakeEngine:
    nl = 12

    def loss_and_grads(self, tokens, prefix, train_gpt=True, mean_reduce=False, on_layer_done=None, before_backward=None):
        LOG.append(("kernels", "forward+head"))
        for l in reversed(range(self.nl)):
            LOG.append(("kernels", f"block{l}"))
            on_layer_done(l)
        LOG.append(("kernels", "embed+mapper"))


def test_segment_capture_and_replay_order(monkeypatch):
    from capdec_b200 import trainer as T
    LOG.clear(); FakeGraph.n = 0; FakeGraph.live = []
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: CUR[-1])
    monkeypatch.setattr(torch.cuda, "graph_pool_handle", lambda: (0, 0))
    monkeypatch.setattr(T.ops, "step_clock", lambda *a, **k: LOG.append(("kernels", "clock")))
    monkeypatch.setattr(torch.distributed, "all_reduce",
                        lambda t, group=None: LOG.append(("all_reduce", int(t[0]), CUR[-1].name,
                                                          any(g.open for g in FakeGraph.live))))
    tr = object.__new__(T.Trainer)
    tr.eng = FakeEngine()
    tr.eng.seed = None
    tr.dev = "cpu"
    tr.step_dev = tr.lr_dev = tr.t_dev = None
    tr.lr, tr.warmup, tr.total = 1e-3, 1, 10
    tr.noise_variance = 0.0
    tr.prefix_d = tr.tokens_d = None
    tr.train_gpt, tr.overlap, tr.segmented, tr.opt_overlap, tr.pg = True, False, True, False, None
    tr.peer = tr.push = False
    tr.buckets = [torch.tensor([float(l)]) for l in range(12)]      # bucket tag = layer index
    tr.head_bucket = torch.tensor([99.0])
    tr.layer_spans = [(100 + l, 101 + l) for l in range(12)]        # parameter span tag = 100 + layer index
    tr.head_span = (199, 200)
    tr.tail = torch.tensor([-7.0, 0.0, 0.0, 0.0])                   # [n_valid, loss_sum, ., .]: tag -7
    tr.stats = torch.zeros(4)
    tr.comm = FakeStream("comm")
    monkeypatch.setattr(T.Trainer, "_adamw_span", lambda self, lo, hi: LOG.append(("adamw", lo, CUR[-1].name)))

    segs = tr._capture_segments()
    assert len(segs) == 13
    assert [int(b[0]) for _, b, _ in segs] == [11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 99]
    assert [sp[0] for _, _, sp in segs] == [111, 110, 109, 108, 107, 106, 105, 104, 103, 102, 101, 100, 199]
    # captures are strictly sequential, each closed before the next opens, and no collective happened while capturing
    caps = [e for e in LOG if e[0] in ("begin", "end")]
    assert caps == [x for i in range(1, 14) for x in (("begin", i), ("end", i))]
    assert not [e for e in LOG if e[0] == "all_reduce"]
    # segment 1 holds clock + forward + head + block 11; the last one the embedding scatter + mapper backward
    first_end = LOG.index(("end", 1))
    assert ("kernels", "forward+head") in LOG[:first_end] and ("kernels", "block11") in LOG[:first_end]
    assert LOG.index(("kernels", "embed+mapper")) > LOG.index(("begin", 13))

    LOG.clear()
    tr._segs = segs
    tr._replay_segments()
    order = [e for e in LOG if e[0] in ("replay", "all_reduce", "adamw")]
    expect = []
    for i, (g, b, sp) in enumerate(segs):
        expect += [("replay", g.id)]
        if i == 0:
            expect += [("all_reduce", -7, "comm", False)]           # the global token count, before any update
        expect += [("all_reduce", int(b[0]), "comm", False), ("adamw", sp[0], "comm")]
    assert order == expect
    # every collective waits for an event recorded on the main stream after its segment; main joins the side stream last
    assert LOG.count(("wait_event", "comm", "main")) == 13
    assert LOG[0] == ("wait_stream", "comm", "main")                # updates must not overtake the previous step's forward
    assert LOG[-1] == ("wait_stream", "main", "comm")
