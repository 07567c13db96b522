"""bench.py's output contract, checked on the CPU with the reference arm (`--impl reference` times the reference's own extractor on
the host cores and needs no GPU): stdout carries exactly one line, that line is JSON, and it has the keys the driver reads.  Native
libraries that print to stdout (NCCL's version banner) must not be able to break this, so bench.py points file descriptor 1 at
stderr for the run and writes the result to the saved descriptor."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.splitlines()
    assert len(lines) == 1, p.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("batched ORB extraction, 640x480") and d["config"]["width"] == 640 and "@640x480" in d["metric"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_stdout_written_by_native_code_is_kept_off_the_result_stream():
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; bench.capture_stdout(); "
            "os.write(1, b'NCCL version 0.0.0\\n'); print('python noise'); bench.emit({'ok': 1})" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout == json.dumps({"ok": 1}) + "\n"
    assert "NCCL version" in p.stderr and "python noise" in p.stderr
