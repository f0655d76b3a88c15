import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def repo_root():
  return ROOT


@pytest.fixture(scope="session", autouse=True)
def built_library():
  """The C-ABI library is a build artefact (git-ignored): build it once if this checkout has none.
  nvcc cross-compiles sm_100a without a GPU, so this works on the CPU-only box too."""
  from surjectors_b200 import _abi

  if not os.path.exists(_abi.LIB_PATH):
    import __graft_entry__ as entry

    entry.build()
  return _abi.LIB_PATH
