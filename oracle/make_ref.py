#!/usr/bin/env python
"""Test/benchmark infrastructure, not product code.

Copy the UNMODIFIED reference hot path into the git-ignored `baseline/_ref/` so that it
travels to the GPU box with the gpurun snapshot (`/root/reference` only exists in the
build container).  Nothing is edited: the files are byte copies, listed with their
SHA-256 in `baseline/_ref/MANIFEST.json`.

    python oracle/make_ref.py            (run by __graft_entry__.build() when /root/reference is present)

What is copied (SURVEY.md 8c caveat 5):
    models/*.py  configs/*.py  tool/*.py  test.py  verify.py  train.py  ckpt/*.pth

`load_reference()` imports it (from baseline/_ref, else from /root/reference) the way
tests/golden/make_golden.py does: an empty `zstandard` stub (tool/process_data.py:16 imports
it, the hot path never calls it) and the reference root in front of sys.path so that
`verify.py`'s `from test import model_test` finds the reference's test.py, not the stdlib's.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')
PARTS = ('models', 'configs', 'tool', 'ckpt')
FILES = ('test.py', 'verify.py', 'train.py')


def copy_reference(src: str = SRC, dst: str = DST) -> dict:
    """Byte-copy the listed files; returns {relative path: sha256}."""
    manifest = {}
    os.makedirs(dst, exist_ok=True)

    def put(rel):
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, 'rb') as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    for part in PARTS:
        for name in sorted(os.listdir(os.path.join(src, part))):
            if name.endswith(('.py', '.pth')):
                put(os.path.join(part, name))
    for name in FILES:
        put(name)
    with open(os.path.join(dst, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'files': manifest}, f, indent=1, sort_keys=True)
    return manifest


def reference_root():
    """Directory holding the unmodified reference, or None."""
    for cand in (DST, SRC):
        if os.path.isfile(os.path.join(cand, 'models', 'user_model.py')):
            return cand
    return None


def verify_copy(dst: str = DST) -> bool:
    """True when every file of baseline/_ref still has the recorded checksum (i.e. is unmodified)."""
    path = os.path.join(dst, 'MANIFEST.json')
    if not os.path.isfile(path):
        return False
    with open(path) as f:
        files = json.load(f)['files']
    for rel, digest in files.items():
        try:
            with open(os.path.join(dst, rel), 'rb') as f:
                if hashlib.sha256(f.read()).hexdigest() != digest:
                    return False
        except OSError:
            return False
    return True


_SHADOWED = ('models', 'configs', 'tool', 'test', 'verify')


class reference_modules:
    """Context manager: inside it `models.*`, `configs.*`, `tool.*`, `test`, `verify` resolve to the UNMODIFIED
    reference; the previous sys.path / sys.modules entries (our shim, the stdlib `test` package) come back on exit.

        with reference_modules() as ref:
            UserModel = ref.UserModel; ref.model_test(...); ref.model_validation(...); ref.auc_score(...)
    """

    def __init__(self, with_scripts: bool = True):
        self.root = reference_root()
        self.with_scripts = with_scripts

    def __enter__(self):
        if self.root is None:
            raise RuntimeError('the reference is neither at baseline/_ref (python oracle/make_ref.py) nor at /root/reference')
        self._saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in _SHADOWED}
        for k in self._saved:
            del sys.modules[k]
        self._path = list(sys.path)
        self._dont = sys.dont_write_bytecode
        sys.dont_write_bytecode = True                    # /root/reference is read-only
        sys.path.insert(0, self.root)
        sys.modules.setdefault('zstandard', types.ModuleType('zstandard'))
        ns = types.SimpleNamespace(root=self.root)
        from models.user_model import UserModel           # noqa: E402  (the reference's)
        ns.UserModel = UserModel
        if self.with_scripts:
            import test as ref_test
            import verify as ref_verify
            from tool.evaluation import auc_score
            ns.model_test, ns.model_validation, ns.auc_score = ref_test.model_test, ref_verify.model_validation, auc_score
            ns.test, ns.verify = ref_test, ref_verify
        assert os.path.abspath(sys.modules['models.user_model'].__file__).startswith(os.path.abspath(self.root))
        return ns

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k.split('.')[0] in _SHADOWED]:
            del sys.modules[k]
        sys.modules.update(self._saved)
        sys.path[:] = self._path
        sys.dont_write_bytecode = self._dont
        return False


if __name__ == '__main__':
    if not os.path.isdir(SRC):
        raise SystemExit(f'{SRC} is not mounted here; baseline/_ref is only refreshed in the build container')
    m = copy_reference()
    print(f'copied {len(m)} files from {SRC} to {DST}')
