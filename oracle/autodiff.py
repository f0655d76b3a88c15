"""value_and_grad and optax-style Adam over the oracle.  TEST INFRASTRUCTURE.

The reference trains with jax.value_and_grad(loss_fn) where loss = -sum(lp)
(examples/coupling_inference_surjection.py:81-90) and optax.adam(1e-4)
(optax 0.2.4 defaults b1=0.9, b2=0.999, eps=1e-8, eps_root=0, bias-corrected).
"""

import torch


def value_and_grad(model, params, reduction="sum", **batch):
  """Return (loss, grads) with loss = -sum(lp) (or -mean) and grads shaped like params."""
  leaves = {m: {k: v.detach().clone().requires_grad_(True) for k, v in b.items()} for m, b in params.items()}
  lp = model.apply(leaves, **batch)
  loss = -(lp.sum() if reduction == "sum" else lp.mean())
  flat = [v for b in leaves.values() for v in b.values()]
  grads = torch.autograd.grad(loss, flat, allow_unused=True)
  it = iter(grads)
  out = {}
  for m, b in leaves.items():
    out[m] = {}
    for k, v in b.items():
      g = next(it)
      out[m][k] = torch.zeros_like(v) if g is None else g
  return loss.detach(), out, lp.detach()


def adam_init(params):
  zeros = lambda: {m: {k: torch.zeros_like(v) for k, v in b.items()} for m, b in params.items()}  # noqa: E731
  return {"count": 0, "mu": zeros(), "nu": zeros()}


def adam_update(params, grads, state, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
  """optax.adam: mu,nu EMA; bias correction 1-b^t; update = -lr * mu_hat / (sqrt(nu_hat) + eps)."""
  t = state["count"] + 1
  new_p, mu, nu = {}, {}, {}
  for m in params:
    new_p[m], mu[m], nu[m] = {}, {}, {}
    for k in params[m]:
      g = grads[m][k]
      mu[m][k] = b1 * state["mu"][m][k] + (1 - b1) * g
      nu[m][k] = b2 * state["nu"][m][k] + (1 - b2) * g * g
      mu_hat = mu[m][k] / (1 - b1**t)
      nu_hat = nu[m][k] / (1 - b2**t)
      new_p[m][k] = params[m][k] - lr * mu_hat / (torch.sqrt(nu_hat) + eps)
  return new_p, {"count": t, "mu": mu, "nu": nu}
