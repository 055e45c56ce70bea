"""Elliptical slice sampling helpers of the PCA-subspace sampler (reference util.py:260-354, re-exported by
``ursabench_b200.util`` under the reference's names): the slice update itself is host logic -- a handful of scalars per
proposal -- and follows the reference's RNG call order so that its chains are reproduced bit for bit
(tests/golden/ess.npz); ``log_pdf`` keeps the projection and the loss on the device."""
import math

import numpy as np
import torch

__all__ = ["cross_entropy", "log_pdf", "elliptical_slice"]


def cross_entropy(model, input, target):
    """(loss, output, {}) -- the criterion signature ``log_pdf`` expects (reference util.py:277-284)."""
    output = model(input)
    return torch.nn.functional.cross_entropy(output, target), output, {}


def log_pdf(theta, subspace, model, loader, criterion, temperature, device):
    """-(sum over the loader of batch loss x batch size) / temperature at the weights ``subspace(theta)``, model in train mode
    (reference util.py:260-274).  The projection runs on ``device`` (``SubspaceModel`` = one K2b launch) and lands in the
    parameters without a host round trip; the loss is accumulated on the device and read back once."""
    device = torch.device(device)
    t = theta if torch.is_tensor(theta) else torch.as_tensor(np.asarray(theta, dtype=np.float32))
    w = subspace(t.to(device=device, dtype=torch.float32))
    offset = 0
    for param in model.parameters():
        param.data.copy_(w[offset:offset + param.numel()].view(param.size()))
        offset += param.numel()
    model.train()
    with torch.no_grad():
        loss = torch.zeros((), device=device)
        for data, target in loader:
            data, target = data.to(device, non_blocking=True), target.to(device, non_blocking=True)
            batch_loss, _, _ = criterion(model, data, target)
            loss += batch_loss * data.size()[0]
    return -loss.item() / temperature


def elliptical_slice(initial_theta, prior, lnpdf, cur_lnpdf=None, angle_range=None, subspace=None, **kwargs):
    """Markov-chain update for a density with a Gaussian prior factored out (Murray, Adams & MacKay 2010); argument
    meaning, RNG call order and return value of reference util.py:287-354 (host logic: a handful of scalars per proposal;
    the cost is ``lnpdf``)."""
    D = len(initial_theta)
    if cur_lnpdf is None:
        cur_lnpdf = lnpdf(initial_theta, subspace, **kwargs)
    if len(prior.shape) == 1:                                  # a sample from the prior
        nu = prior
    else:                                                      # chol(Sigma)
        if not prior.shape[0] == D or not prior.shape[1] == D:
            raise IOError("Prior must be given by a D-element sample or DxD chol(Sigma)")
        nu = np.dot(prior, np.random.normal(size=D))
    hh = math.log(np.random.uniform()) + cur_lnpdf             # slice threshold
    if angle_range is None or angle_range == 0.:
        phi = np.random.uniform() * 2. * math.pi               # whole ellipse, both bracket edges at the first proposal
        phi_min = phi - 2. * math.pi
        phi_max = phi
    else:
        phi_min = -angle_range * np.random.uniform()
        phi_max = phi_min + angle_range
        phi = np.random.uniform() * (phi_max - phi_min) + phi_min
    while True:
        xx_prop = initial_theta * math.cos(phi) + nu * math.sin(phi)
        cur_lnpdf = lnpdf(xx_prop, subspace, **kwargs)
        if cur_lnpdf > hh:
            break
        if phi > 0:
            phi_max = phi
        elif phi < 0:
            phi_min = phi
        else:
            raise RuntimeError("BUG DETECTED: Shrunk to current position and still not acceptable.")
        phi = np.random.uniform() * (phi_max - phi_min) + phi_min
    return (xx_prop, cur_lnpdf)
