"""Host logic of the BRICK adaptor layer without a GPU: B200StreamBatcher (sora_b200/brick/b200_bricks.hpp) collects the windows of K graph
instances, lets every instance stage its own window in one shared arena and hands the round to the engine as one call.  The C ABI is stubbed
(sora_b200/brick/batcher_selftest.cpp); the stub's "frame" is a checksum of the samples it received, so a region that overlaps another or holds
stale data fails the run.  Also built with ThreadSanitizer when the toolchain has it."""
import os, subprocess, shutil, pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "sora_b200", "brick", "batcher_selftest.cpp")

def _build(tmp_path, extra):
    exe = str(tmp_path / ("selftest" + ("_tsan" if extra else "")))
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-pthread"] + extra + ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "sora_b200", "brick"), "-o", exe, SRC])
    return exe

@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
@pytest.mark.parametrize("threads,rounds", [(1, 100), (3, 400), (16, 150)])
def test_batcher_rounds_and_arena(tmp_path, threads, rounds):
    exe = _build(tmp_path, [])
    out = subprocess.run([exe, str(threads), str(rounds)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bad 0" in out.stdout, out.stdout + out.stderr

@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_batcher_is_race_free_under_thread_sanitizer(tmp_path):
    try: exe = _build(tmp_path, ["-fsanitize=thread"])
    except subprocess.CalledProcessError: pytest.skip("ThreadSanitizer runtime not available")
    out = subprocess.run([exe, "8", "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bad 0" in out.stdout and "WARNING: ThreadSanitizer" not in out.stderr, out.stdout + out.stderr[-2000:]
