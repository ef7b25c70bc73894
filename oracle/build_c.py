"""Builds the C part of the oracle (oracle/nle_oracle_c.c -> oracle/libnle_oracle_c.so) with gcc.
TEST INFRASTRUCTURE ONLY.  -ffp-contract=off keeps the reference's evaluation order (no FMA contraction)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "nle_oracle_c.c")
LIB = os.path.join(HERE, "libnle_oracle_c.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
           SRC, "-o", LIB, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed building libnle_oracle_c.so")
    return LIB


if __name__ == "__main__":
    print(build(force=True))
