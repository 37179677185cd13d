"""Static SASS opcode histogram of the hot kernels of libtiler_slider.so (cuobjdump -sass), so that
the instruction counts DESIGN.md steers by are reproducible offline:
    python profiles/experiments/sass_histogram.py > profiles/r2_sass_histogram.json
Per kernel: static instruction count, opcode classes, global loads/stores, and -- for straight-line
kernels -- static instructions per env-step (step kernels process 4 envs per thread, wide 1)."""
import json
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "tiler_slider_b200", "libtiler_slider.so")
KERNELS = {  # mangled-name fragment -> (label, envs per thread or None)
    "step_kernelILi6ELi4ELi0ELb1ELi1E": ("step_kernel<6,4,ordered,auto-reset,u8 count> (c3)", 4),
    "step_kernelILi5ELi1ELi0ELb1ELi1E": ("step_kernel<5,1,ordered,auto-reset,u8 count> (c2)", 4),
    "wide_step_kernelILi8ELi0ELb1ELi6E": ("wide_step_kernel<8,ordered,auto-reset,LW=6> (c4)", 1),
    "valid_kernelILi6ELi4E": ("valid_kernel<6,4>", 4),
    "goal_kernelILi6ELi4E": ("goal_kernel<6,4>", 4),
    "observe_kernelILi6E": ("observe_kernel<6>", None),
    "bfs_local_kernelILi6ELi4E": ("bfs_local_kernel<6,4> (K6)", None),
    "bfs_expand_kernelILi6ELi4E": ("bfs_expand_kernel<6,4> (K4)", None),
    "bfs_expand_exchange_kernelILi6ELi4E": ("bfs_expand_exchange_kernel<6,4> (K4x)", None),
    "bfs_hash_insert_kernel": ("bfs_hash_insert_kernel (K5)", None),
    "generic_step_kernel": ("generic_step_kernel", None),
}
CLASSES = {"LOP3": "logic", "SHF": "shift", "PRMT": "permute", "SEL": "select", "IMAD": "imad", "IADD3": "iadd", "POPC": "popc",
           "BREV": "brev", "LDG": "ld.global", "STG": "st.global", "LDS": "ld.shared", "STS": "st.shared", "ATOMS": "atom.shared",
           "ATOMG": "atom.global", "RED": "red.global", "ISETP": "setp", "BRA": "branch", "BAR": "barrier", "LDC": "ld.const",
           "S2R": "sreg", "VOTE": "vote", "SHFL": "shuffle", "LEA": "lea", "MOV": "mov", "FSEL": "select", "UBLKCP": "bulk-copy (TMA 1-D)"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out, cur, counts = {}, None, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = next((k for k in KERNELS if k in m.group(1)), None)
            counts = Counter() if cur else None
            if cur:
                out[cur] = counts
            continue
        if counts is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            counts[m.group(1)] += 1
    res = {}
    for k, c in out.items():
        label, per = KERNELS[k]
        total = sum(v for op, v in c.items() if op != "NOP")
        cls = Counter()
        for op, v in c.items():
            cls[CLASSES.get(op, op.lower())] += v
        res[label] = {"static_instructions": total, "nop_padding": c.get("NOP", 0),
                      "static_instructions_per_env_step": (total / per if per else None),
                      "classes": dict(sorted(cls.items(), key=lambda kv: -kv[1])[:16])}
    json.dump({"library": os.path.relpath(LIB, ROOT), "how": "cuobjdump -sass, opcodes counted per function; "
               "straight-line kernels (the tail/auto-reset branches included), so static counts bound the dynamic ones from above",
               "kernels": res}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
