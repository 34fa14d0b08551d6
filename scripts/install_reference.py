#!/usr/bin/env python
"""Install the UNMODIFIED reference (mhpi/hydrodl2) into the git-ignored `baseline/_ref/`.

    python scripts/install_reference.py [/root/reference]

`bench.py --impl reference` and the `cpu_baseline` leg time the reference's own PyTorch path
(`hydrodl2.load_model('hbv', 'Hbv')` -> `forward` -> `backward`) from this directory when it
exists; otherwise they fall back to the oracle port (bit-exact against it, tests/).  The
directory is git-ignored but travels to the GPU box with `gpurun` (see .gitignore /
.gpurunignore).  No reference source enters the repository's history.

1. the documented way:
       pip install --no-index --no-build-isolation --find-links /opt/wheelhouse \
           --target baseline/_ref <copy of the reference>
   needs the reference's build backend (hatchling + hatch-vcs, pyproject.toml:1-3), which is
   neither installed nor in the offline wheelhouse -> fails here ("No module named 'hatchling'").
2. fallback: the package is pure Python (no extension modules), so a wheel is its
   `src/hydrodl2` tree plus the `_version.py` that hatch-vcs would generate
   (`hydrodl2/__init__.py:11,17` refuses to import without it) — that is what is laid down.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, 'baseline', '_ref')


def install(src_root: str = '/root/reference', quiet: bool = False) -> str:
    """-> 'pip' | 'copy' | 'absent' (reference not on this machine)."""
    if not os.path.isdir(os.path.join(src_root, 'src', 'hydrodl2')):
        return 'absent'
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix='hydroref_') as tmp:
        work = os.path.join(tmp, 'reference')
        shutil.copytree(src_root, work)            # /root/reference is read-only; builds write into the tree
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps',
               '--find-links', '/opt/wheelhouse', '--target', DEST, work]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode == 0 and os.path.isdir(os.path.join(DEST, 'hydrodl2')):
            how = 'pip'
        else:
            if not quiet:
                tail = (r.stdout or '').strip().splitlines()[-1:] or ['']
                print(f'pip install failed ({tail[0]}); laying down the pure-Python package instead')
            shutil.copytree(os.path.join(work, 'src', 'hydrodl2'), os.path.join(DEST, 'hydrodl2'))
            with open(os.path.join(DEST, 'hydrodl2', '_version.py'), 'w') as f:
                f.write("__version__ = '1.0.0+ref'\n")
            how = 'copy'
    with open(os.path.join(DEST, 'INSTALL_NOTE.txt'), 'w') as f:
        f.write(f'unmodified mhpi/hydrodl2 from {src_root}, installed by scripts/install_reference.py ({how})\n')
    return how


if __name__ == '__main__':
    print(install(sys.argv[1] if len(sys.argv) > 1 else '/root/reference'), DEST)
