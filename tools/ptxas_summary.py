"""Summarise `nvcc -Xptxas -v` logs (one per translation unit) into a table: kernel, registers, static shared memory, spill
stores / loads, stack frame.  Usage: python tools/ptxas_summary.py LOG [LOG ...]   (kernel names demangled with c++filt)"""
import re
import subprocess
import sys


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*$", "", o) for o in out]


def main(paths):
    rows = []
    for p in paths:
        cur = None
        for line in open(p, errors="replace"):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur = {"unit": p.rsplit("/", 1)[-1].replace(".log", ".cu"), "name": m.group(1), "stack": 0, "spill_st": 0, "spill_ld": 0, "regs": 0, "smem": 0}
                rows.append(cur)
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                cur["stack"], cur["spill_st"], cur["spill_ld"] = (int(v) for v in m.groups())
            m = re.search(r"Used (\d+) registers", line)
            if m:
                cur["regs"] = int(m.group(1))
                s = re.search(r"(\d+) bytes smem", line)
                cur["smem"] = int(s.group(1)) if s else 0
    names = demangle([r["name"] for r in rows])
    print("%-14s %-72s %5s %8s %8s %8s %6s" % ("unit", "kernel", "regs", "smem B", "spill st", "spill ld", "stack"))
    for r, n in sorted(zip(rows, names), key=lambda t: (t[0]["unit"], t[1])):
        print("%-14s %-72s %5d %8d %8d %8d %6d" % (r["unit"], n[:72], r["regs"], r["smem"], r["spill_st"], r["spill_ld"], r["stack"]))
    spilled = [n for r, n in zip(rows, names) if r["spill_st"] or r["spill_ld"]]
    print("\n%d kernels, %d with register spills%s" % (len(rows), len(spilled), (": " + ", ".join(sorted(set(spilled)))) if spilled else ""))


if __name__ == "__main__":
    main(sys.argv[1:])
