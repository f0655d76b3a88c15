"""as_batch_iterator (surjectors/util.py:50-103): batch arithmetic against the oracle on the CPU; the on-device row gather
(surj_index_take) bit-exact against numpy.take on the GPU."""

from collections import namedtuple

import numpy as np
import pytest
import torch

from oracle import util as outil
from surjectors_b200 import _abi, util


@pytest.mark.parametrize("n,bs", [(1000, 100), (1001, 100), (99, 100), (100, 100), (1, 7), (4097, 4096)])
def test_batch_layout_matches_oracle(n, bs):
  assert util.batch_layout(n, bs) == outil.batch_layout(n, bs)
  nb, b = util.batch_layout(n, bs)
  assert (nb - 1) * b < n <= nb * b


def test_cpu_tensors_are_refused():
  data = namedtuple("data", "y")(torch.zeros(10, 2))
  with pytest.raises(_abi.SurjError):
    util.as_batch_iterator(0, data, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("n,shape_y,shape_x,bs,shuffle", [
  (1000, (2,), None, 100, True),        # the reference's docstring example
  (1003, (20,), (16,), 100, True),      # ragged last batch, conditional data set, float4 path
  (517, (3,), (5,), 64, True),          # widths that are not multiples of 4: scalar path
  (300, (4, 6), None, 128, False),      # trailing dims beyond one, no shuffle
  (50, (8,), None, 100, True),          # n < batch_size: one batch of n
])
def test_gather_bit_exact(n, shape_y, shape_x, bs, shuffle):
  rng = np.random.Generator(np.random.PCG64(n))
  y = rng.standard_normal((n,) + shape_y).astype(np.float32)
  fields = {"y": y}
  if shape_x is not None:
    fields["x"] = rng.standard_normal((n,) + shape_x).astype(np.float32)
  nt = namedtuple("data", " ".join(fields))
  data = nt(*[torch.from_numpy(a).cuda() for a in fields.values()])
  it = util.as_batch_iterator(7, data, bs, shuffle)
  nb, b = outil.batch_layout(n, bs)
  assert it.num_batches == nb
  idxs = it.idxs.cpu().numpy()
  assert idxs.dtype == np.int32 and sorted(idxs.tolist()) == list(range(n))
  if not shuffle:
    assert (idxs == np.arange(n)).all()
  seen = 0
  for j in range(it.num_batches):
    got = it(j)
    want = outil.get_batch(fields, idxs, j, b)
    assert set(got) == set(want)
    for k in want:
      g = got[k].cpu().numpy()
      assert g.shape == want[k].shape
      assert (g.view(np.uint32) == want[k].view(np.uint32)).all()  # bit-exact
    seen += len(want["y"])
  assert seen == n
  assert int(it.error_flag.item()) == 0
  with pytest.raises(IndexError):
    it(it.num_batches)


@pytest.mark.gpu
def test_user_indices_and_out_of_range_flag():
  y = torch.arange(40, dtype=torch.float32, device="cuda").reshape(10, 4)
  it = util.as_batch_iterator(0, namedtuple("data", "y")(y), 5, shuffle=False)
  mine = torch.tensor([9, 8, 7, 6, 5, 4, 3, 2, 1, 0], dtype=torch.int32, device="cuda")
  got = it(1, mine)["y"].cpu().numpy()
  assert (got == y.cpu().numpy()[[4, 3, 2, 1, 0]]).all()
  bad = torch.tensor([0, 1, 99, 3, 4, 5, 6, 7, 8, 9], dtype=torch.int32, device="cuda")
  got = it(0, bad)["y"].cpu().numpy()
  assert (got[2] == 0).all() and (got[1] == y.cpu().numpy()[1]).all()
  assert int(it.error_flag.item()) == 1
