import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'render_golden.npz'))


@pytest.fixture(scope='session')
def viewless_golden():
    """Outputs of the unmodified reference with use_viewdirs=False (oracle/make_golden_viewless.py) and the state-dicts it ran."""
    import numpy as np
    import torch
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'viewless_golden.npz'))
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    sds = []
    for pre, tag in (('coarse/', 'c'), ('fine/', 'f')):
        sd = {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre + 'pts_linears.')}
        sd['views_linears.0.weight'] = torch.from_numpy(z[pre + 'views_linears.0.weight'][:, :256].copy())
        sd['views_linears.0.bias'] = torch.from_numpy(z[pre + 'views_linears.0.bias'])
        sd['output_linear.weight'] = torch.from_numpy(g[tag + '_output_w'])
        sd['output_linear.bias'] = torch.from_numpy(g[tag + '_output_b'])
        sds.append(sd)
    return g, tuple(sds)


@pytest.fixture(scope='session')
def wfit():
    """(sd_coarse, sd_fine): the analytic-scene weights made by oracle/make_weights.py."""
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    sdc = {k[len('coarse/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('coarse/')}
    sdf = {k[len('fine/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
    return sdc, sdf
