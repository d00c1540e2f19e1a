"""Static check of the rule DESIGN.md §3.1 came out of round 2 with -- a shared-memory stage is handed back to its producer only
after instructions that CONSUME what was read from it -- on the FP64 kernels, which recycle their TMA stages through a
shared-memory counter (ptx::stage_release: MEMBAR + ATOMS): in the SASS of every kernel, the last LDS before the first ATOMS must
be followed, before that ATOMS, by arithmetic (DMMA / DFMA / ...) -- i.e. the loads cannot still be in flight when the release
issues (in-order issue).  Usage: python profiles/release_order_audit.py > profiles/release_order_r02.txt"""
import glob
import os
import re
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sorted(glob.glob(os.path.join(root, "build", "obj", "rb_*.o")) + glob.glob(os.path.join(root, "build", "obj", "pass_np*.o")) +
              [os.path.join(root, "build", "obj", "jade.o")])
consumers = ("DMMA", "DFMA", "DMUL", "DADD")
print("# " + __doc__.strip().replace("\n", "\n# "))
print(f"{'kernels':>8} {'min consumers between last LDS and release':>44}  object")
worst_overall = None
for o in objs:
    sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True, check=True).stdout
    n_k, worst = 0, None
    for f in sass.split("Function : ")[1:]:
        ops = [m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)]
        atoms = [i for i, op in enumerate(ops) if op.startswith("ATOMS")]
        if not atoms:
            continue
        a = atoms[0]
        lds = [i for i, op in enumerate(ops) if op.startswith("LDS") and i < a]
        if not lds:
            continue
        n_k += 1
        n_cons = sum(1 for op in ops[lds[-1] + 1:a] if op.startswith(consumers))
        worst = n_cons if worst is None else min(worst, n_cons)
    if n_k:
        print(f"{n_k:8d} {worst:44d}  {os.path.basename(o)}")
        worst_overall = worst if worst_overall is None else min(worst_overall, worst)
print(f"# fewest arithmetic instructions between the last shared-memory load and the release, over all kernels: {worst_overall}")
