"""Static check of the built library (no GPU needed): in every kernel that takes part in a
programmatic-dependent-launch chain, no global load may be scheduled in front of the
griddepcontrol.wait (SASS: ACQBULK).  nvcc turns loads through `const T* __restrict__` into
ld.global.nc and hoists them freely; one that moved above the wait read a count the predecessor
kernel had not written yet (k_gftt_pick, found on a B200)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "esvio_b200", "csrc", "libesvio_fe.so")


def _cuobjdump():
    for c in (shutil.which("cuobjdump"), "/usr/local/cuda/bin/cuobjdump"):
        if c and os.path.exists(c):
            return c
    return None


@pytest.mark.skipif(_cuobjdump() is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_no_global_load_in_front_of_the_pdl_wait():
    sass = subprocess.run([_cuobjdump(), "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, before, done, n_chain, bad = None, [], False, 0, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn, before, done = m.group(1), [], False
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not m or fn is None or done:
            continue
        op = m.group(1).strip()
        if "ACQBULK" in op:
            done = True
            n_chain += 1
            loads = [o for o in before if re.search(r"(^|\s)(LDG|LD|ATOMG|ATOM|REDG?)[.\s]", o)]
            if loads:
                bad[fn] = loads
        else:
            before.append(op)
    assert n_chain >= 15, n_chain          # the kernels of a window's chain were found
    assert not bad, bad
