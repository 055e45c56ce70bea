"""On-disk sample-bank format: one file per ensemble + reference-style per-sample state_dict export
(the reference's interchange is ``torch.save(model.state_dict(), 'sghmc_sample_%d.pt')``, experiment.py:77-80, read back
by trtprof/run_prediction.py:50-57).  CPU-only: the format code never touches a kernel."""
import copy

import pytest
import torch
import torch.nn as nn

from ursabench_b200.bank import SampleBank


def _net():
    return nn.Sequential(nn.Conv2d(3, 5, 3, padding=1, bias=False), nn.BatchNorm2d(5), nn.ReLU(),
                         nn.Flatten(), nn.Linear(5 * 4 * 4, 7))


def _ensemble(n, seed=0):
    torch.manual_seed(seed)
    models = []
    for _ in range(n):
        m = _net()
        m[1].running_mean.normal_()
        m[1].running_var.uniform_(0.5, 2.0)
        models.append(m.eval())
    return models


def _bank(models):
    bank = SampleBank.from_modules(models, "cpu")
    bank.skeleton = copy.deepcopy(models[0])
    return bank


def test_bank_file_round_trip(tmp_path):
    models = _ensemble(3)
    bank = _bank(models)
    assert bank.D == sum(p.numel() for p in models[0].parameters()) and bank.nb == 10
    path = str(tmp_path / "ens.bank.pt")
    bank.save(path)
    back = SampleBank.load(path, "cpu", skeleton=_net())
    assert back.count == 3 and back.D == bank.D and back.nb == bank.nb
    assert torch.equal(back.w[:3, :bank.D], bank.w[:3, :bank.D]) and torch.equal(back.b[:3, :10], bank.b[:3, :10])
    x = torch.randn(4, 3, 4, 4)
    for h, m in zip(back.handles(), models):
        assert torch.equal(h.eval()(x), m(x))          # lazily materialised sample == the original module, bit for bit


def test_bank_layout_descriptor_and_mismatch(tmp_path):
    bank = _bank(_ensemble(1))
    lay = bank.layout()
    assert [e["name"] for e in lay["params"]] == [n for n, _ in _net().named_parameters()]
    assert [e["name"] for e in lay["buffers"]] == ["1.running_mean", "1.running_var"]
    assert lay["params"][1]["offset"] == 3 * 5 * 9
    path = str(tmp_path / "b.pt")
    bank.save(path)
    other = nn.Sequential(nn.Conv2d(3, 5, 3, padding=1, bias=True), nn.BatchNorm2d(5), nn.ReLU(), nn.Flatten(), nn.Linear(80, 7))
    with pytest.raises(ValueError):
        SampleBank.load(path, "cpu", skeleton=other)
    torch.save({"format": "something else"}, path)
    with pytest.raises(ValueError):
        SampleBank.load(path, "cpu")


def test_reference_style_state_dict_export_import(tmp_path):
    models = _ensemble(2, seed=3)
    bank = _bank(models)
    paths = bank.export_state_dicts(str(tmp_path / "sghmc_sample_%d.pt"))
    assert len(paths) == 2
    x = torch.randn(2, 3, 4, 4)
    for path, m in zip(paths, models):
        fresh = _net().eval()
        fresh.load_state_dict(torch.load(path))          # exactly what trtprof/run_prediction.py:55-56 does
        assert torch.equal(fresh(x), m(x))
        assert set(torch.load(path).keys()) == set(m.state_dict().keys())
    again = SampleBank.from_state_dict_files(paths, _net(), "cpu")
    assert again.count == 2 and torch.equal(again.w[:2, :bank.D], bank.w[:2, :bank.D])
    assert torch.equal(again.b[:2, :bank.nb], bank.b[:2, :bank.nb])


def test_empty_bank_round_trip(tmp_path):
    bank = SampleBank(12, 0, "cpu", capacity=1)
    path = str(tmp_path / "e.pt")
    bank.save(path)
    back = SampleBank.load(path, "cpu")
    assert back.count == 0 and back.D == 12 and back.nb == 0 and back.handles() == []
