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
    """The tail of bench.py's own arm with the three legs stubbed: rank 0 prints one line carrying all legs; other ranks
    print nothing; --no-* flags leave nulls; cpu_baseline only at N=1."""
    body = (
        "import argparse\n"
        "calls = []\n"
        "bench.cpu_reference = lambda **k: calls.append('cpu') or {'value': 700.0, 'cores': 16, 'windows': 64, 'steps_done': 10, 'host_cores': 128}\n"
        "bench.spmm_probe = lambda dev, pk: calls.append('spmm') or {'achieved': 1500.0}\n"
        "bench.train_probe = lambda *a: calls.append('train') or {'value': 59000.0}\n"
        "def run(rank, world, **flags):\n"
        "    args = argparse.Namespace(no_cpu=False, no_spmm=False, no_train=False); args.__dict__.update(flags)\n"
        "    line = {'value': 1.0, 'spmm': None, 'train': None, 'cpu_baseline': None}\n"
        "    bench._secondary_legs(line, args, rank, world, None, {}, None, None, None)\n"
        "    return line\n"
        "a = run(0, 1); assert calls == ['cpu', 'spmm', 'train'], calls\n"
        "assert a['cpu_baseline']['value'] == 700.0 and a['cpu_baseline']['kind'] == 'port' and a['spmm'] and a['train']\n"
        "calls.clear(); b = run(1, 2); assert calls == ['train'] and b['cpu_baseline'] is None and b['spmm'] is None\n"
        "calls.clear(); c = run(0, 2); assert calls == ['spmm', 'train'] and c['cpu_baseline'] is None\n"
        "calls.clear(); d = run(0, 1, no_cpu=True, no_spmm=True, no_train=True); assert calls == [] and d['train'] is None\n"
    )
    r = _run(body)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 3                                      # the three rank-0 calls; rank 1 printed nothing
    first = json.loads(lines[0])
    assert first["cpu_baseline"]["cores"] == 16 and "legs_skipped" not in first
