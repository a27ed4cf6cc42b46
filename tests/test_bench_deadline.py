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
