"""Recipe that puts the UNMODIFIED reference on the GPU box (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference's hot path is a pure-Python tree: nothing to compile.  `/root/reference`
exists only in the build container, so `build()` copies -- byte for byte, no edits -- the files the renderer, the
TriPlaneGenerator around it and their import chain need into ``oracle/_ref/g_nerf/`` (git-ignored, NOT
gpurun-ignored: it travels to the GPU box like a built .so).  Nothing under ``oracle/_ref`` is ever committed and no
reference source is copied anywhere else in the repository.

Who may use it: ``tests/`` (same-device parity against the reference itself), ``bench.py --impl reference`` (the CPU arm),
``bench.py``'s ``gpu_baseline`` / ``config3`` legs (the reference renderer and the reference's backbone + super-resolution
on the same B200) and ``__graft_entry__.smoke()``.  The product (``g-nerf_b200/``) never imports it; the launcher is
handed a reference checkout by path exactly as a user would hand it one.

    python oracle/build_ref.py          # here: copy from /root/reference
    oracle.build_ref.reference_dir()    # -> path of a usable g_nerf directory, or None
"""
import filecmp
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = '/root/reference/g_nerf'
OUT_ROOT = os.path.join(HERE, '_ref', 'g_nerf')
MANIFEST = os.path.join(HERE, '_ref', 'MANIFEST.json')

# relative to g_nerf/: the hot path (VR/), its callers (triplane.py) and what those import at module level
TREES = ['training/volumetric_rendering', 'torch_utils', 'dnnlib']
FILES = ['camera_utils.py', 'legacy.py', 'gen_videos.py', 'training/__init__.py', 'training/triplane.py',
         'training/networks_stylegan2.py', 'training/networks_stylegan3.py', 'training/superresolution.py',
         'training/base_network.py', 'training/audio_network.py', 'training/crosssection_utils.py']
SKIP_DIRS = {'__pycache__'}


def _walk(tree):
    base = os.path.join(SRC_ROOT, tree)
    for d, dirs, files in os.walk(base):
        dirs[:] = [x for x in dirs if x not in SKIP_DIRS]
        for f in files:
            if not f.endswith('.pyc'):
                yield os.path.relpath(os.path.join(d, f), SRC_ROOT)


def _sha(path):
    h = hashlib.sha256()
    with open(path, 'rb') as fh:
        h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False):
    """Copy the subtree when the reference checkout is present (the build container); on the GPU box the copy made
    here is used as is.  Returns the g_nerf directory or None when neither exists."""
    if not os.path.isdir(SRC_ROOT):
        return OUT_ROOT if os.path.isdir(OUT_ROOT) else None
    rels = sorted(set(FILES + [r for t in TREES for r in _walk(t)]))
    manifest = {}
    for rel in rels:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(OUT_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(MANIFEST, 'w') as fh:
        json.dump({'source': SRC_ROOT, 'files': manifest}, fh, indent=1, sort_keys=True)
    return OUT_ROOT


def verify():
    """True when every file of the manifest is present and unmodified (sha256)."""
    if not os.path.exists(MANIFEST):
        return False
    files = json.load(open(MANIFEST))['files']
    return all(os.path.exists(os.path.join(OUT_ROOT, r)) and _sha(os.path.join(OUT_ROOT, r)) == h for r, h in files.items())


def reference_dir():
    """A g_nerf directory holding the unmodified reference: the travelling copy if it exists, else the checkout."""
    if os.path.isdir(os.path.join(OUT_ROOT, 'training', 'volumetric_rendering')):
        return OUT_ROOT
    if os.path.isdir(os.path.join(SRC_ROOT, 'training', 'volumetric_rendering')):
        return SRC_ROOT
    return None


if __name__ == '__main__':
    print(build(force=True), 'verified' if verify() else 'NOT verified')
