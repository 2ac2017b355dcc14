#!/usr/bin/env python
"""Copy the UNMODIFIED reference files of the hot path into the git-ignored baseline/_ref/ (it travels to the GPU box with the
gpurun snapshot; /root/reference does not exist there).

`pip install /root/reference` is not possible: the reference has no setup.py / pyproject.toml (it is a directory of scripts),
so the "install" is this byte-for-byte copy of the six modules the path needs, with their package layout (`utils/`, `models/`).
bench.py imports them from baseline/_ref for `--impl reference` (CPU) and for the `gpu_eager_baseline` leg (same modules with
device='cuda:0').  Nothing under dyglib_b200/ imports them.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ['utils/utils.py', 'utils/DataLoader.py', 'models/modules.py', 'models/TGAT.py', 'models/DyGFormer.py',
         'models/MemoryModel.py']


def install(reference='/root/reference', dest=None, quiet=False):
    dest = dest or os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(reference, 'models')):
        if not quiet:
            print(f'{reference} not present: nothing installed (baseline/_ref is used as it is)')
        return False
    lines = []
    for rel in FILES:
        out = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(reference, rel), out)
        lines.append(f'{hashlib.sha256(open(out, "rb").read()).hexdigest()}  {rel}')
    with open(os.path.join(dest, 'SHA256SUMS'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    if not quiet:
        print('\n'.join(lines))
    return True


if __name__ == '__main__':
    sys.exit(0 if install(*sys.argv[1:2]) else 1)
