"""The C-ABI library builds, loads, and exports every entry point include/nsr_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built_lib():
    import neural_sim_nerf_b200.build as b
    return b.build()


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'nsr_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nsr_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ('nsr_pack_net', 'nsr_mlp_forward', 'nsr_raw2outputs', 'nsr_sample_pdf', 'nsr_resample_merge',
              'nsr_render_rays_forward', 'nsr_make_rays', 'nsr_last_error'):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    h = ctypes.CDLL(built_lib)
    for s in declared_symbols():
        assert hasattr(h, s), f'{s} declared in include/nsr_b200.h but not exported'


def test_python_binding_covers_header(built_lib):
    import neural_sim_nerf_b200 as nsr
    assert sorted(nsr.EXPORTED_SYMBOLS) == declared_symbols()
    L = nsr.lib()
    assert L.nsr_version() >= 100
    # 73 forward + 78 backward (transposed) + 73 mixed-precision forward operand chunk pairs (32 KiB each) + fp32 tail
    # 73 forward + 78 backward + 73 mixed-precision chunk pairs, the fp32 tail, the fp32 transposed density branch (refinement)
    assert L.nsr_packed_net_bytes() == (73 + 78 + 73) * 32768 + 3360 * 4 + (256 * (63 + 256 * 4 + 319 + 256 * 2) + 256) * 4
    assert L.nsr_render_backward_workspace_bytes(512, 192) >= 512 * 192 * (16 + 32)
    assert L.nsr_render_workspace_bytes(0, 64, 128) == 0
    assert L.nsr_render_workspace_bytes(512, 64, 128) >= 512 * (64 * 4 * 2 + 64 * 16 + 192 * 4 + 192 * 16)


def test_parameter_errors_do_not_need_a_gpu(built_lib):
    import neural_sim_nerf_b200 as nsr
    L = nsr.lib()
    rc = L.nsr_mlp_forward(None, None, 4, 64, None, 0, None, None)
    assert rc == -1 and b'null' in L.nsr_last_error()
    rc = L.nsr_raw2outputs(None, None, None, 3, -1, 64, 0, None, None, None, None, None, None)
    assert rc == -1
    assert L.nsr_mlp_forward(None, None, 0, 64, None, 0, None, None) == 0   # empty batch is a no-op


def test_sass_is_blackwell_native(built_lib):
    """tcgen05.mma / tcgen05.ld / bulk-copy must be in the binary (UTCHMMA / LDTM / UBLKCP)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', built_lib], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTCQMMA', 'LDTM', 'UBLKCP'):
        assert mnemonic in sass, mnemonic


def test_chunk_issue_order_invariants(built_lib):
    """The schedule both MLP kernels rely on (DESIGN.md "Pipeline"): every (half, K chunk) of a step exactly once; in a two-half
    step all chunks whose A operand is ready early -- an encoding, or activations 0..127 -- come before any late one (so the first
    operand half may be overwritten as soon as accumulator 0 is complete), and accumulator 0's last chunk precedes accumulator 1's."""
    import neural_sim_nerf_b200 as nsr
    L = nsr.lib()
    k_chunks = [1, 4, 4, 4, 4, 5, 4, 4, 4, 5]
    halves = [2] * 9 + [1]
    early = [1, 2, 2, 2, 2, 3, 2, 2, 2, None]
    total = 0
    for step in range(10):
        h = (ctypes.c_int * 16)()
        k = (ctypes.c_int * 16)()
        n = L.nsr_chunk_issue_order(step, h, k, 16)
        assert n == k_chunks[step] * halves[step]
        total += n
        slots = [(h[i], k[i]) for i in range(n)]
        assert sorted(slots) == [(a, b) for a in range(halves[step]) for b in range(k_chunks[step])]
        if halves[step] == 2:
            is_late = [kc >= early[step] for _, kc in slots]
            assert is_late == sorted(is_late), (step, slots)            # every early chunk before every late one
            last0 = max(i for i, (a, _) in enumerate(slots) if a == 0)
            last1 = max(i for i, (a, _) in enumerate(slots) if a == 1)
            assert last0 < last1 == n - 1
            first_late = is_late.index(True) if True in is_late else n
            assert all(a == 0 for a, _ in slots[:early[step]]) and all(a == 1 for a, _ in slots[early[step]:2 * early[step]])
            assert first_late == 2 * early[step] or first_late == n
        else:
            assert slots == [(0, b) for b in range(k_chunks[step])]
    assert total == 73                                                     # NUM_CHUNKS
    assert L.nsr_chunk_issue_order(10, (ctypes.c_int * 16)(), (ctypes.c_int * 16)(), 16) == -1
    assert L.nsr_chunk_issue_order(1, (ctypes.c_int * 16)(), (ctypes.c_int * 16)(), 3) == -1
