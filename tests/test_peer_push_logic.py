"""CPU dry run of the data-parallel step's HOST ordering in the peer / push variant (Trainer._fwd_bwd, _push_span, _peer_opt):
torch.cuda's stream / event objects, the engine, the C-ABI calls and torch.distributed are replaced by recording fakes.
Checked: the gradient buffer is cleared on the side stream before the first gradient write is allowed; every finished
GPT-2 block is pushed (copy_async to the peers that own it) on the push stream after an event on the main stream, in the
order backward finishes the blocks, [mapper | wte | wpe] last; the step ends with the main stream joining the push stream;
the update is count all-reduce -> one kernel with local gradient slices -> fence all-reduce, and it does not clear the
gradients (the next step's forward does).  The numerical check runs on 2 GPUs (tests/test_dp_gpu.py)."""
import contextlib

import torch

LOG = []


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
        self.on = (stream or CUR[-1]).name


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
    nl = 3
    seed = None

    def loss_and_grads(self, tokens, prefix, train_gpt=True, mean_reduce=False, on_layer_done=None, before_backward=None):
        LOG.append(("kernels", "forward+head", CUR[-1].name))
        before_backward()
        LOG.append(("kernels", "lm_head_wgrad", CUR[-1].name))
        for l in reversed(range(self.nl)):
            LOG.append(("kernels", f"block{l}", CUR[-1].name))
            on_layer_done(l)
        LOG.append(("kernels", "embed+mapper", CUR[-1].name))


class FakeFlat:
    """stands in for the flat gradient / parameter views: only data_ptr() is used on this path"""
    def __init__(self, base):
        self.base = base

    def data_ptr(self):
        return self.base


def test_push_variant_host_ordering(monkeypatch):
    from capdec_b200 import trainer as T
    LOG.clear()
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: CUR[-1])
    monkeypatch.setattr(T.ops, "step_clock", lambda *a, **k: LOG.append(("kernels", "clock", CUR[-1].name)))
    monkeypatch.setattr(T.ops, "zero_fill", lambda t: LOG.append(("zero_fill", CUR[-1].name)))
    monkeypatch.setattr(T.ops, "copy_async", lambda dst, src, n: LOG.append(("copy", CUR[-1].name, dst, src, n)))
    monkeypatch.setattr(T.ops, "adamw_peer_step", lambda g, p, rank, lo, n, *a, **k: LOG.append(("peer_kernel", tuple(g), rank, lo, n)))
    monkeypatch.setattr(torch.distributed, "all_reduce", lambda t, group=None: LOG.append(("all_reduce", t.tag, CUR[-1].name)))

    world, rank, shard = 2, 1, 40
    tr = object.__new__(T.Trainer)
    tr.eng = FakeEngine()
    tr.dev = "cpu"
    tr.step_dev = tr.lr_dev = tr.t_dev = None
    tr.lr, tr.warmup, tr.total = 1e-3, 1, 10
    tr.noise_variance = 0.0
    tr.prefix_d = tr.tokens_d = None
    tr.train_gpt, tr.overlap, tr.segmented, tr.opt_overlap, tr.pg = True, False, False, False, None
    tr.peer = tr.push = True
    tr.world, tr.rank, tr.shard = world, rank, (rank * shard, (rank + 1) * shard)
    tr.betas, tr.eps, tr.wd = (0.9, 0.999), 1e-6, 0.0
    tr.m_flat = tr.v_flat = None
    # 80 trainable "parameters": [head 0..20 | block0 20..40 | block1 40..60 | block2 60..80]; rank 0 owns 0..40, rank 1 owns 40..80
    tr.head_span = (0, 20)
    tr.layer_spans = [(20, 40), (40, 60), (60, 80)]
    tr.g_flat = FakeFlat(1000)
    tr.staging_ptrs = [5000, 9000]            # staging areas of rank 0 (a peer address) and of this rank
    tr.push_stream, tr.zero_stream = FakeStream("push"), FakeStream("zero")
    tr.g_slices, tr.p_ptrs = [9000, 1000 + 4 * 40], [7000, 8000]

    class Tagged:
        def __init__(self, tag):
            self.tag = tag

        def copy_(self, other):
            LOG.append(("copy_", self.tag, other.tag))

        def __getitem__(self, sl):
            return self
    tr.stats, tr.tail, tr.fence = Tagged("stats"), Tagged("tail"), Tagged("fence")

    tr._fwd_bwd()
    names = [e for e in LOG if e[0] in ("zero_fill", "kernels", "copy", "wait_event", "wait_stream")]
    # the clear runs on the zero stream, after an event of the main stream, and the first gradient write waits for it
    i_zero = names.index(("zero_fill", "zero"))
    assert ("wait_event", "zero", "main") in names[:i_zero]
    i_join = names.index(("wait_event", "main", "zero"))
    assert i_zero < i_join < names.index(("kernels", "lm_head_wgrad", "main"))
    assert names.index(("kernels", "forward+head", "main")) < i_join          # the forward pass does not wait for the clear
    # this rank (1) owns 40..80, so only what rank 0 owns is pushed: block 0 (20..40) and the head (0..20), into slot 0 of
    # rank 0's staging area (this rank is the only other source), at the element's offset inside rank 0's slice
    copies = [e for e in LOG if e[0] == "copy"]
    assert copies == [("copy", "push", 5000 + 4 * 20, 1000 + 4 * 20, 4 * 20), ("copy", "push", 5000, 1000, 4 * 20)]
    # each push follows the kernels that finished its bucket, through an event recorded on the main stream
    i_b0 = names.index(("kernels", "block0", "main"))
    i_c0 = names.index(copies[0])
    assert i_b0 < i_c0 and ("wait_event", "push", "main") in names[i_b0:i_c0]
    assert names.index(("kernels", "embed+mapper", "main")) < names.index(copies[1])
    assert names[-1] == ("wait_stream", "main", "push")                          # the step ends with every push issued

    LOG.clear()
    tr._peer_opt()
    assert LOG == [("copy_", "stats", "tail"), ("all_reduce", "stats", "main"),
                   ("peer_kernel", (9000, 1000 + 4 * 40), 1, 40, 40), ("all_reduce", "fence", "main")]
