"""Dev aid (no GPU needed): generate + NVRTC-compile the fused kernel of a BP operator and report registers / SASS mix.
   usage: CEED_B200_COMPILE_ONLY=1 python scripts/dev_compile.py BP P [EPB]"""
import os, subprocess, sys, tempfile
os.environ.setdefault("CEED_B200_COMPILE_ONLY", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bp, p = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3: os.environ["CEED_B200_EPB"] = sys.argv[3]
d = tempfile.mkdtemp(prefix="b200jit_")
os.environ["CEED_B200_DUMP_CUBIN"] = d
from libceed_b200 import Ceed
from libceed_b200.bp import BPProblem
ceed = Ceed()
prob = BPProblem(ceed, bp, p, (4, 4, 4), build_qdata=False)
src = prob.op.kernel_source()
info = prob.op.kernel_info()
print("plan:", info)
cubins = sorted(f for f in os.listdir(d) if f.endswith(".cubin"))
cubin = os.path.join(d, cubins[-1])
res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout
print(res.strip())
sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
import collections, re
ops = collections.Counter(m.group(1).split(".")[0] for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass, re.M))
print({k: v for k, v in ops.most_common(25)})
print("source:", os.path.join(d, cubins[-1].replace(".cubin", ".cu")), "lines:", src.count("\n"))
