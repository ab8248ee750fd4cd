"""Run an UNMODIFIED reference script (gen_videos.py, ...) with the B200 renderer dropped in.

    python -m gnerf_b200.launch --reference /path/to/G-NeRF/g_nerf [--no-plane-cache] [--decoder-precision bf16]
                                [--channels-first] gen_videos.py --network ... --id_encoder ... --id_image ...

What it does before handing over to the script with ``runpy`` ("ship install() plus a launcher"):
  1. puts the reference checkout on ``sys.path`` and calls ``install()`` (class-level patch of the reference's
     ImportanceRenderer / RaySampler / MipRayMarcher2, g-nerf_b200/install.py);
  2. makes the reference's own CUDA plugins loadable under a current PyTorch (``enable_reference_plugins``);
  3. wraps ``legacy.load_network_pkl`` so that every TriPlaneGenerator it returns gets the backbone / repacked-plane
     cache (frames.enable_plane_cache: gen_videos.py re-runs the backbone for each of its 120 frames with the same ws,
     gen_videos.py:150,171) and the renderer options chosen on the command line in its ``rendering_kwargs``.
The reference's files are not modified and nothing is copied from them.
"""
import argparse
import importlib
import os
import runpy
import sys

from . import frames
from .install import install


def enable_reference_plugins():
    """Let the reference's OWN CUDA plugins (bias_act / upfirdn2d, used by its backbone and super-resolution head, which stay on
    the reference path) load under a current PyTorch.  The reference's loader calls ``torch.utils.cpp_extension.load(name=...)``
    and then ``importlib.import_module(name)`` (torch_utils/custom_ops.py:141-144): that relied on ``load`` leaving the built
    module in ``sys.modules``, which PyTorch 2.x no longer does (it returns the module without registering it), so every
    plugin fails with ModuleNotFoundError right after compiling.  This wraps ``load`` to register what it returns -- the
    behaviour the reference was written against.  The reference's files are not touched.  Idempotent."""
    import torch.utils.cpp_extension as ext
    if getattr(ext.load, '_tpr_registers_module', False):
        return
    original = ext.load

    def load(name, *args, **kwargs):
        module = original(name, *args, **kwargs)
        if kwargs.get('is_python_module', True) and hasattr(module, '__name__'):
            sys.modules.setdefault(name, module)
        return module

    load._tpr_registers_module = True
    load.__wrapped__ = original
    ext.load = load


def _configure_generators(obj, extra_options, plane_cache):
    """Apply the launcher's settings to every TriPlaneGenerator-shaped module in a loaded pickle dict."""
    done = []
    items = obj.items() if isinstance(obj, dict) else []
    for key, net in items:
        if all(hasattr(net, a) for a in ('renderer', 'backbone', 'decoder', 'rendering_kwargs')):
            net.rendering_kwargs.update(extra_options)
            if plane_cache:
                frames.enable_plane_cache(net)
            done.append(key)
    return done


def patch_loader(legacy_module, extra_options, plane_cache=True):
    """Wrap ``legacy.load_network_pkl`` (legacy.py:17) in place; returns the original."""
    original = legacy_module.load_network_pkl
    if getattr(original, '_tpr_wrapped', False):
        return original

    def load_network_pkl(*args, **kwargs):
        data = original(*args, **kwargs)
        _configure_generators(data, extra_options, plane_cache)
        return data

    load_network_pkl._tpr_wrapped = True
    load_network_pkl.__wrapped__ = original
    legacy_module.load_network_pkl = load_network_pkl
    return original


def main(argv=None):
    ap = argparse.ArgumentParser(prog='gnerf_b200.launch', description=__doc__.split('\n\n')[0])
    ap.add_argument('--reference', required=True, help="the reference checkout's g_nerf directory")
    ap.add_argument('--no-plane-cache', action='store_true', help='re-run backbone and repack on every synthesis call')
    ap.add_argument('--decoder-precision', default='fp32', choices=['fp32', 'bf16', 'fp32_ffma'])
    ap.add_argument('--channels-first', action='store_true',
                    help='write the feature image as [N,32,H,W] so the permute+contiguous at triplane.py:81 is a no-op')
    ap.add_argument('script', help='reference script to run, relative to --reference or absolute')
    ap.add_argument('script_args', nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)

    ref = os.path.abspath(args.reference)
    if not os.path.isdir(os.path.join(ref, 'training', 'volumetric_rendering')):
        raise SystemExit(f'{ref} does not look like the reference g_nerf directory (no training/volumetric_rendering)')
    if ref not in sys.path:
        sys.path.insert(0, ref)
    install()
    enable_reference_plugins()
    extra = {'decoder_precision': args.decoder_precision}
    if args.channels_first:
        extra['output_layout'] = 'channels_first'
    patch_loader(importlib.import_module('legacy'), extra, plane_cache=not args.no_plane_cache)
    script = args.script if os.path.isabs(args.script) else os.path.join(ref, args.script)
    sys.argv = [script] + list(args.script_args)
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
