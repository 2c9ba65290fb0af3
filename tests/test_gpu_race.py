"""Concurrent Add / Delete / Update / Read / Search / DestroySet through the C ABI (tools/race_test.cpp), plain and
under ThreadSanitizer.  The reference races on exactly these paths (set.go:77-116 appends to s.Blocks while Search
ranges over it, cache.go:104-115 walks AllBlocks without the mutex, block.go's IDs[] is read by Search while Insert
writes it); SURVEY section 5 asks for race tooling around the replacement's locks.

The TSAN variant links a second build of the library whose HOST code is instrumented (tools/Makefile
`race_test_tsan`, built by __graft_entry__.build()); it is skipped when that build is not there."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tools")


def test_concurrent_mutations_and_searches(built):
    exe = os.path.join(TOOLS, "race_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", TOOLS, "race_test"], stdout=subprocess.DEVNULL)
    out = subprocess.run([exe, "250"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "race_test ok" in out.stdout, out.stdout[-500:] + out.stderr[-500:]


def test_thread_sanitizer_is_silent(built):
    exe = os.path.join(TOOLS, "race_test_tsan")
    if not os.path.exists(exe):
        pytest.skip("tools/race_test_tsan not built (make -C tools race_test_tsan)")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66 report_signal_unsafe=0")
    out = subprocess.run([exe, "120"], capture_output=True, text=True, timeout=600, env=env)
    assert "race_test ok" in out.stdout, out.stdout[-500:] + out.stderr[-1500:]
    assert out.returncode == 0 and "WARNING: ThreadSanitizer" not in out.stderr, out.stderr[-3000:]
