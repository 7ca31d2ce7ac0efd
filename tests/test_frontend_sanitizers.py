"""Race and memory checking of the circuit front-end (the reference has no race detection of its own; the
front-end here emits one witness on several host threads, and bench.py runs eight such passes at once).  The driver
tests/sanitize/frontend_sanitize.cpp is compiled together with frontend/circuits.cpp under ThreadSanitizer and under
AddressSanitizer + UndefinedBehaviorSanitizer and must exit 0 with no report: concurrent provers, 2-4 threads per pass, dirty and
reused destination buffers, every cell compared with the one-thread pass; the keygen pass; two error paths.  CPU only."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "tests", "sanitize", "frontend_sanitize.cpp"),
       os.path.join(ROOT, "delay-encryption-in-halo2_b200", "frontend", "circuits.cpp")]
OUT = os.path.join(ROOT, "tests", "sanitize")


def _sanitizer_usable(flag: str, tmp_path) -> bool:
    probe = tmp_path / "probe.cpp"
    probe.write_text("#include <thread>\nint x; int main() { std::thread t([] { x = 1; }); t.join(); return x - 1; }\n")
    exe = tmp_path / "probe"
    if subprocess.run(["g++", "-O1", flag, "-pthread", str(probe), "-o", str(exe)], capture_output=True).returncode != 0:
        return False
    return subprocess.run([str(exe)], capture_output=True, timeout=60).returncode == 0


@pytest.mark.parametrize("name,flags", [("tsan", ["-fsanitize=thread"]),
                                        ("asan_ubsan", ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"])])
def test_frontend_under_sanitizer(name, flags, tmp_path):
    if not _sanitizer_usable(flags[0], tmp_path):
        pytest.skip(f"{flags[0]} runtime not usable on this machine")
    exe = os.path.join(OUT, f"frontend_{name}")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-pthread"] + flags + SRC + ["-o", exe])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66", ASAN_OPTIONS="detect_leaks=1 exitcode=67", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, f"exit {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-6000:]}"
    assert r.stdout.strip().endswith("ok")
    assert "WARNING: ThreadSanitizer" not in r.stderr and "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr
