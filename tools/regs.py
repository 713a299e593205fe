"""Register / spill listing of the kernel unit: python tools/regs.py [filter ...] (runs nvcc -Xptxas -v on csrc/mrc_b200.cu)."""
import os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
defs = [a for a in sys.argv[1:] if a.startswith("-D")]
sys.argv = [a for a in sys.argv if not a.startswith("-D")]
cmd = ["nvcc", *defs, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xptxas", "-v", "-Xcompiler", "-fPIC",
       "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "min_ratio_cycle_b200", "csrc"),
       "-c", os.path.join(ROOT, "min_ratio_cycle_b200", "csrc", "mrc_b200.cu"), "-o", "/tmp/mrc_regs.o"]
t = subprocess.run(cmd, capture_output=True, text=True)
log = t.stdout + t.stderr
if t.returncode:
    sys.exit(log)
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                 r".*Used (\d+) registers")
fpat = re.compile(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads")
seen = set()
for m in fpat.finditer(log):   # device functions called through the ABI (noinline)
    if m.group(1) in seen or int(m.group(3)) + int(m.group(4)) == 0:
        continue
    seen.add(m.group(1))
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    if "(" in name and not name.startswith("void ") or (len(sys.argv) > 1 and not any(f in name for f in sys.argv[1:])):
        continue
    if re.search(r"Compiling entry function '" + re.escape(m.group(1)) + "'", log):
        continue
    print(f"  fn {name[:84]:84s} stack {m.group(2):>4s}  spill st/ld {m.group(3)}/{m.group(4)}")
for m in pat.finditer(log):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    if len(sys.argv) > 1 and not any(f in name for f in sys.argv[1:]):
        continue
    print(f"{name[:90]:90s} regs {m.group(5):>3s}  stack {m.group(2):>4s}  spill st/ld {m.group(3)}/{m.group(4)}")
