"""CPU: the C-ABI library loads, exports every symbol include/triplane_b200.h declares, and rejects
bad arguments before touching a device (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'triplane_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tpr_[a-z_0-9]+)\s*\(', text)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for s in ('tpr_render', 'tpr_run_model', 'tpr_ray_sample', 'tpr_ray_march', 'tpr_sample_pdf',
              'tpr_sample_importance', 'tpr_decode', 'tpr_pack_planes', 'tpr_pack_decoder', 'tpr_last_error'):
        assert s in syms


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.lib()
    for s in header_symbols():
        assert hasattr(lib, s), f'{s} declared in the header but not exported'
    assert set(header_symbols()) == set(pkg._lib.EXPORTED_SYMBOLS), 'binding and header disagree'
    assert lib.tpr_abi_version() == 8


def test_options_struct_layout_matches_header(pkg):
    # 3 doubles + 10 int32 (= 64) + density_noise double + its two draw pointers
    assert ctypes.sizeof(pkg._lib.TprOptions) == 88


def test_peer_sinks_struct_layout_matches_header(pkg):
    # 2 int32 + 3 arrays of 15 pointers
    assert ctypes.sizeof(pkg._lib.TprPeerSinks) == 8 + 3 * 15 * 8
    lib = pkg._lib.lib()
    assert lib.tpr_peer_alloc(0, None, None) == -1 and lib.tpr_peer_open(None, None) == -1
    assert lib.tpr_peer_close(None) == -1 and lib.tpr_peer_free(None) == -1


def test_argument_errors_are_reported_without_a_device(pkg):
    lib = pkg._lib.lib()
    assert lib.tpr_pack_planes(None, 1, 8, 8, None, None) == -1
    assert b'NULL' in lib.tpr_last_error()
    assert lib.tpr_packed_planes_bytes(2, 256, 256) == 2 * 3 * 32 * 256 * 256 * 4
    assert lib.tpr_packed_decoder_bytes() == 4452 * 4
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.tpr_ray_sample(p, p, 0, 4, p, p, None) == -2              # n_img <= 0
    assert lib.tpr_sample_pdf(p, 3, p, p, 1, 5, 4, p, None, None) == -2  # bins_stride < n_weights + 1
    o = pkg._lib.TprOptions(ray_start=2.25, ray_end=3.3, box_warp=1.0, depth_resolution=200,
                            depth_resolution_importance=200)
    rc = lib.tpr_render(p, 1, 8, 8, p, p, p, 4, p, p, None, None, ctypes.byref(o), p, p, p, None, None, None, 1, p, 256, None)
    assert rc == -2 and b'depth resolutions' in lib.tpr_last_error()
    with pytest.raises(RuntimeError, match='code -2'):
        pkg._lib.check(rc, 'tpr_render')
    o = pkg._lib.TprOptions(ray_start=2.25, ray_end=3.3, box_warp=1.0, depth_resolution=8, depth_resolution_importance=8,
                            plane_sets=3)
    rc = lib.tpr_render(p, 4, 8, 8, p, p, p, 4, p, p, None, None, ctypes.byref(o), p, p, p, None, None, None, 1, p, 1024, None)
    assert rc == -2 and b'plane_sets' in lib.tpr_last_error()            # 4 images over 3 plane sets
    o = pkg._lib.TprOptions(ray_start=2.25, ray_end=3.3, box_warp=1.0, depth_resolution=8, depth_resolution_importance=8,
                            output_layout=7)
    rc = lib.tpr_render(p, 4, 8, 8, p, p, p, 4, p, p, None, None, ctypes.byref(o), p, p, p, None, None, None, 1, p, 1024, None)
    assert rc == -3 and b'output_layout' in lib.tpr_last_error()
    o = pkg._lib.TprOptions(depth_resolution=8, depth_clamp_group=2)
    assert lib.tpr_render_scratch_bytes(8, 16, ctypes.byref(o)) == 512 + 8 * 4   # one (min, max) pair per clamp slot
    assert lib.tpr_sample_stratified(None, 4, None, None, ctypes.byref(o), p, None) == -1
    assert lib.tpr_sample_stratified(p, 4, p, None, ctypes.byref(o), p, None) == -1  # per-ray limits need both


def test_host_shim_rejects_cpu_tensors_and_bad_decoders(pkg):
    import torch
    R = pkg.ImportanceRenderer()
    opts = {'ray_start': 2.25, 'ray_end': 3.3, 'box_warp': 1, 'depth_resolution': 8,
            'depth_resolution_importance': 8, 'disparity_space_sampling': False, 'clamp_mode': 'softplus'}
    planes = torch.zeros(1, 3, 32, 8, 8)
    dec = pkg.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
    rays = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError, match='no CPU path'):
        R(planes, dec, rays, rays, opts)
    with pytest.raises(RuntimeError, match='no CPU path'):
        pkg.RaySampler()(torch.eye(4)[None], torch.eye(3)[None], 4)
    with pytest.raises(RuntimeError, match='OSGDecoder-shaped'):
        pkg.pack_decoder(torch.nn.Linear(32, 33))
    assert sorted(dec.state_dict()) == ['net.0.bias', 'net.0.weight', 'net.2.bias', 'net.2.weight']
    assert R.plane_axes.shape == (3, 3, 3) and len(list(R.parameters())) == 0


def test_reference_module_names_are_importable(pkg):
    """every public name of the reference's renderer module exists here too."""
    from importlib import import_module
    r = import_module('g-nerf_b200.volumetric_rendering.renderer')
    for name in ('generate_planes', 'project_onto_planes', 'sample_from_planes', 'sample_from_3dgrid', 'ImportanceRenderer'):
        assert callable(getattr(r, name)), name
    for name in ('forward', 'run_model', 'sort_samples', 'unify_samples', 'sample_stratified', 'sample_importance', 'sample_pdf'):
        assert callable(getattr(r.ImportanceRenderer, name)), name
    import torch
    axes = r.generate_planes()
    xyz = torch.tensor([[[0.1, 0.2, 0.3]]])
    got = r.project_onto_planes(axes, xyz)                       # plane 0 <- (x,y), 1 <- (x,z), 2 <- (z,x)
    torch.testing.assert_close(got, torch.tensor([[[0.1, 0.2]], [[0.1, 0.3]], [[0.3, 0.1]]]))
    with pytest.raises(RuntimeError, match='no CPU path'):
        r.sample_from_planes(axes, torch.zeros(1, 3, 32, 4, 4), xyz, box_warp=1)


def test_measurement_library_is_separate_from_the_product(pkg):
    """The microbenchmarks and the raw tcgen05 layer test live in libtriplane_b200_bench.so (include/triplane_b200_bench.h);
    the product library exports none of them."""
    text = open(os.path.join(ROOT, 'include', 'triplane_b200_bench.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    declared = sorted(set(re.findall(r'\b(tpr_[a-z_0-9]+)\s*\(', text)))
    assert set(declared) == set(pkg._lib.BENCH_EXPORTED_SYMBOLS)
    bench, prod = pkg._lib.bench_lib(), pkg._lib.lib()
    for s in declared:
        assert hasattr(bench, s) and not hasattr(prod, s), s
