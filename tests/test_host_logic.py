"""CPU: host-side logic around the C engine that needs no GPU -- the parameter-sync bookkeeping (per-key and batched
upload), and the ordering contract of bench.py's training configuration under torchrun."""
import ast
import os

import torch

from helping_hand_for_egocentric_videos_b200.model.LaviLa import _ParamSync

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _params():
    return [("a.weight", torch.nn.Parameter(torch.randn(4, 3))), ("a.bias", torch.nn.Parameter(torch.randn(4))),
            ("b.weight", torch.nn.Parameter(torch.randn(2, 4).double()))]


def test_param_sync_uploads_only_what_changed():
    ps, named, sent = _ParamSync(), _params(), []
    ps.sync(named, lambda k, t: sent.append((k, t)))
    assert [k for k, _ in sent] == ["a.weight", "a.bias", "b.weight"]
    assert all(t.dtype == torch.float32 and t.is_contiguous() for _, t in sent)   # converted for the fp32 engine
    sent.clear()
    ps.sync(named, lambda k, t: sent.append((k, t)))
    assert sent == []                                                            # nothing changed: nothing re-sent
    with torch.no_grad():
        named[1][1].add_(1.0)                                                    # in-place op bumps the version counter
    ps.sync(named, lambda k, t: sent.append((k, t)))
    assert [k for k, _ in sent] == ["a.bias"]
    sent.clear()
    named[0][1].data.mul_(2.0)                                                   # .data writes do NOT: mark_dirty() = clear()
    ps.sync(named, lambda k, t: sent.append((k, t)))
    assert sent == []
    ps.clear()
    ps.sync(named, lambda k, t: sent.append((k, t)))
    assert len(sent) == 3


def test_param_sync_batched_form_sends_one_list():
    """An optimizer step changes every decoder parameter: the batched setter gets them in ONE call (hh_decoder_set_weights),
    keeps the converted temporaries alive in the list, and is not called at all when nothing changed."""
    ps, named, calls = _ParamSync(), _params(), []
    ps.sync(named, None, batch_setter=lambda items: calls.append(list(items)))
    assert len(calls) == 1 and [k for k, _ in calls[0]] == ["a.weight", "a.bias", "b.weight"]
    assert calls[0][2][1].dtype == torch.float32 and torch.equal(calls[0][2][1], named[2][1].detach().float())
    ps.sync(named, None, batch_setter=lambda items: calls.append(list(items)))
    assert len(calls) == 1
    with torch.no_grad():
        for _, p in named[:2]:
            p.mul_(0.5)
    ps.sync(named, None, batch_setter=lambda items: calls.append(list(items)))
    assert len(calls) == 2 and [k for k, _ in calls[1]] == ["a.weight", "a.bias"]


def test_bench_train_config_steps_on_every_rank_before_rank0_returns():
    """bench.py --config c4 at world > 1: every TrainShare.step() contains the packed all-gather, so no step may be taken
    by rank 0 alone (the profiled step after `if rank != 0: return` once deadlocked the 2-GPU run).  Checked on the source:
    inside run_train no `ts.step(...)` call appears after the first `if rank != 0: return`."""
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_train")
    ret_line = None
    for node in ast.walk(fn):
        if isinstance(node, ast.If) and isinstance(node.test, ast.Compare) and ast.unparse(node.test) == "rank != 0" \
                and any(isinstance(b, ast.Return) for b in node.body):
            ret_line = node.lineno if ret_line is None else min(ret_line, node.lineno)
    assert ret_line is not None
    late = [n.lineno for n in ast.walk(fn) if isinstance(n, ast.Call) and ast.unparse(n.func) == "ts.step" and n.lineno > ret_line]
    assert late == [], "ts.step() after the non-zero ranks returned (lines %s): a collective taken by rank 0 alone" % late
