"""The oracle's frames for a fixed set of scenes, pinned by hash (tests/golden/oracle_frames_sha256.json, written by
tools/make_golden.py).  These are not reference outputs -- the reference's golden PNGs are LFS pointers and it cannot be
built here -- they keep the oracle itself from drifting: the recorded frames are the ones the CUDA path was bit-exact
against (the 32 scenes that already existed at the last commit verified on a B200, 880e5ed, hash the same with that
commit's oracle -- checked when this file was written)."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_frames_match_the_recorded_hashes():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    recorded = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_frames_sha256.json")))
    now = m.generate()
    assert set(now) == set(recorded)
    changed = [name for name in sorted(now) if now[name] != recorded[name]]
    assert not changed, f"the oracle renders these scenes differently from the recorded frames: {changed} (python tools/make_golden.py after a deliberate change)"
