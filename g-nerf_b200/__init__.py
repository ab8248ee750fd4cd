"""g-nerf_b200: a B200-native (sm_100a) tri-plane volume renderer that drops in behind the renderer
API of llrtt/G-NeRF -- ImportanceRenderer / RaySampler / MipRayMarcher2 / OSGDecoder and nothing else.

The directory name contains a hyphen, so import it with
``importlib.import_module('g-nerf_b200')`` or through the ``gnerf_b200`` alias module at the repo root.
"""
from .volumetric_rendering import (ImportanceRenderer, RaySampler, MipRayMarcher2, PackedPlanes,  # noqa: F401
                                   pack_planes, pack_decoder, generate_planes)
from .triplane import OSGDecoder, FullyConnectedLayer  # noqa: F401
from .install import install, uninstall  # noqa: F401
from .build import build  # noqa: F401
from . import parallel  # noqa: F401
from . import frames  # noqa: F401
from .frames import render_frames, synthesize_frames, enable_plane_cache, disable_plane_cache  # noqa: F401
from .launch import enable_reference_plugins  # noqa: F401

__all__ = ['ImportanceRenderer', 'RaySampler', 'MipRayMarcher2', 'OSGDecoder', 'FullyConnectedLayer',
           'PackedPlanes', 'pack_planes', 'pack_decoder', 'generate_planes', 'install', 'uninstall', 'build',
           'render_frames', 'synthesize_frames', 'enable_plane_cache', 'disable_plane_cache', 'enable_reference_plugins']
