"""bench.py must print exactly one JSON line within its wall-clock budget even when a secondary leg (CPU baseline on a
contended host, a probe stuck in native code) does not come back: `_LegDeadline` is exercised in a subprocess."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(body):
    code = "import sys, time; sys.path.insert(0, %r); import bench\n" % ROOT + body
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)


def test_deadline_expires_prints_partial_line_and_exits_zero():
    r = _run("line = {'value': 1.0, 'cpu_baseline': None, 'spmm': None, 'train': None}\n"
             "d = bench._LegDeadline(line, ['cpu_baseline', 'spmm', 'train'], True, 0.5)\n"
             "d.done('cpu_baseline', {'value': 2.0})\n"
             "time.sleep(30)\n"                      # a leg that never comes back
             "print('not reached')\n")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["cpu_baseline"] == {"value": 2.0} and d["train"] is None and d["legs_skipped"]["legs"] == ["spmm", "train"]


def test_deadline_not_reached_prints_once():
    r = _run("line = {'value': 1.0, 'a': None}\n"
             "d = bench._LegDeadline(line, ['a'], True, 1.0)\n"
             "d.done('a', 3)\n"
             "d.finish()\n"
             "time.sleep(1.5)\n"                     # past the (cancelled) deadline: nothing more may be printed
             "d.finish()\n")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"value": 1.0, "a": 3}


def test_non_zero_rank_exits_silently():
    r = _run("d = bench._LegDeadline({'x': None}, ['x'], False, 0.3)\ntime.sleep(30)\n")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_secondary_legs_fill_the_line_in_order_and_print_once():
    """The tail of bench.py's own arm with the legs stubbed: rank 0 prints one line carrying all legs; other ranks
    print nothing; --no-* flags leave nulls; cpu_baseline / reference_gpu only at N=1."""
    body = (
        "import argparse\n"
        "calls = []\n"
        "bench.cpu_reference = lambda **k: calls.append('cpu') or {'value': 700.0, 'cores': 16, 'windows': 64, 'steps_done': 10, 'host_cores': 128}\n"
        "bench.spmm_probe = lambda dev, pk: calls.append('spmm') or {'achieved': 1500.0}\n"
        "bench.reference_gpu_probe = lambda *a: calls.append('refgpu') or {'eager': {'value': 9000.0}}\n"
        "bench.train_probe = lambda *a: calls.append('train') or {'value': 59000.0}\n"
        "hostwin = lambda: calls.append('hostwin') or {'value': 5.0}\n"
        "def run(rank, world, **flags):\n"
        "    args = argparse.Namespace(no_cpu=False, no_spmm=False, no_train=False, no_refgpu=False, no_hostwin=False, windows=64); args.__dict__.update(flags)\n"
        "    line = {'value': 1.0, 'spmm': None, 'train': None, 'cpu_baseline': None}\n"
        "    bench._secondary_legs(line, args, rank, world, None, {}, None, None, None, hostwin)\n"
        "    return line\n"
        "a = run(0, 1); assert calls == ['cpu', 'spmm', 'refgpu', 'hostwin', 'train'], calls\n"
        "assert a['cpu_baseline']['value'] == 700.0 and a['cpu_baseline']['kind'] == 'port' and a['spmm'] and a['train'] and a['reference_gpu']\n"
        "calls.clear(); b = run(1, 2); assert calls == ['hostwin', 'train'] and b['cpu_baseline'] is None and b['spmm'] is None\n"
        "calls.clear(); c = run(0, 2); assert calls == ['spmm', 'hostwin', 'train'] and c['cpu_baseline'] is None and c['reference_gpu'] is None\n"
        "calls.clear(); d = run(0, 1, no_cpu=True, no_spmm=True, no_train=True, no_refgpu=True, no_hostwin=True); assert calls == [] and d['train'] is None\n"
    )
    r = _run(body)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 3                                      # the three rank-0 calls; rank 1 printed nothing
    first = json.loads(lines[0])
    assert first["cpu_baseline"]["cores"] == 16 and "legs_skipped" not in first


def test_failing_leg_is_recorded_and_the_line_still_printed():
    """A secondary leg that raises (the round-1 N=8 failure: a shape error inside the training probe) must not take the
    already-measured headline down: at N=1 it becomes {"error": ...}; under torchrun the line is printed, then it re-raises."""
    body = (
        "import argparse\n"
        "def boom(*a, **k): raise RuntimeError('shape mismatch in probe')\n"
        "bench.cpu_reference = boom; bench.spmm_probe = boom; bench.reference_gpu_probe = boom; bench.train_probe = boom\n"
        "args = argparse.Namespace(no_cpu=False, no_spmm=False, no_train=False, no_refgpu=False, no_hostwin=False, windows=64)\n"
        "line = {'value': 123.0}\n"
        "bench._secondary_legs(line, args, 0, 1, None, {}, None, None, None, boom)\n"
        "line2 = {'value': 456.0}\n"
        "try:\n"
        "    bench._secondary_legs(line2, args, 0, 2, None, {}, None, None, None, boom)\n"
        "except RuntimeError:\n"
        "    print('RERAISED', file=sys.stderr)\n"
    )
    r = _run(body)
    assert r.returncode == 0, r.stderr
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 2 and "RERAISED" in r.stderr
    assert lines[0]["value"] == 123.0 and all("error" in lines[0][k] for k in ("cpu_baseline", "spmm", "reference_gpu", "e2e_host_windows", "train"))
    assert lines[1]["value"] == 456.0 and "shape mismatch" in lines[1]["leg_failure"]


def test_epoch_feeder_serves_full_batches_for_every_world_size():
    """The training probe feeds fixed (64, ...) staging buffers for 9+ steps.  One epoch of a rank's shard holds fewer full
    batches at world 8 (2851 train windows / 8 ranks / 64 = 5) -- the feeder must roll over epochs (set_epoch) and only ever
    hand out full batches (drop_last)."""
    import numpy as np
    import bench

    class Loader(object):   # IndexBatchLoader's host-side contract (len / set_epoch / iter), no GPU
        def __init__(self, n, world, rank, bs):
            from pytorch_geometric_temporal_b200.signal.index_dataset import shard_indices
            self.n, self.world, self.rank, self.bs, self.epoch, self.shard = n, world, rank, bs, 0, shard_indices
        def set_epoch(self, e):
            self.epoch = e
        def __len__(self):
            return len(self.shard(self.n, self.world, self.rank, True, 0, self.epoch)) // self.bs
        def __iter__(self):
            pos = self.shard(self.n, self.world, self.rank, True, 0, self.epoch)
            for s in range(0, len(pos) - self.bs + 1, self.bs):
                yield np.asarray(pos[s:s + self.bs])

    from pytorch_geometric_temporal_b200.signal import index_splits
    tr, _, _ = index_splits(4096, 12)
    for world in (1, 2, 4, 8):
        for rank in (0, world - 1):
            f = bench.EpochFeeder(Loader(len(tr), world, rank, 64))
            seen_epochs = set()
            for _ in range(40):
                b = f.next()
                assert b.shape == (64,)
                seen_epochs.add(f.epoch)
            assert f.batches == 40
            if world == 8:
                assert len(seen_epochs) > 1          # had to roll over
