#!/usr/bin/env python
"""BASELINE config 4 with the 'vs reference' column: tools/sweep.py (GPU MSM / NTT size sweep) with the CPU restatement of the
reference's best_multiexp / best_fft (oracle/, all host cores) timed on the same inputs and every GPU result compared with the
CPU's.  Lives under tests/ because only test infrastructure may touch oracle/.  Not collected by pytest (no test_ prefix).

    python tests/sweep_vs_cpu.py --msm-to 24 --ntt-to 27 --cpu-to 24 > profiles/r02_sweep_with_cpu_reference.jsonl
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import orc  # noqa: E402
import sweep  # noqa: E402

if __name__ == "__main__":
    sweep.main(cpu=orc)
