import io
import json
import os
import sys
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


@pytest.fixture(scope='session')
def datadir(tmp_path_factory):
    """The reference's test inputs (tests/data), unpacked from the fixture."""
    d = tmp_path_factory.mktemp('dn') / 'tests' / 'data'
    d.mkdir(parents=True)
    with tarfile.open(os.path.join(GOLDEN, 'data.tar.gz')) as tf:
        tf.extractall(str(d), filter='data')
    return str(d)


@pytest.fixture(scope='session')
def goldens():
    with open(os.path.join(GOLDEN, 'scan_goldens.json')) as f:
        return json.load(f)
