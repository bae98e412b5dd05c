"""The oracle's restatement of the training step's backward pass against the gradients of the unmodified reference
(tests/golden/train_*.npz, written by make_golden.py from `model.train(); model(x); TrainRLossList; loss.backward()`)."""
import numpy as np
import pytest
import torch

from conftest import TRAIN_CASES, load_golden, oracle_config_for, train_state_dict
from oracle import llicti_oracle as O


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_oracle_gradients_equal_the_reference(name):
    g = load_golden(name)
    ocfg = oracle_config_for(name)
    torch.set_num_threads(4)
    loss, grads = O.train_loss_and_grads(ocfg, train_state_dict(name), g["rgb"])
    assert abs(loss - float(g["loss"])) <= 2e-6 * abs(loss)
    keys = [k[5:] for k in g.files if k.startswith("grad/")]
    assert sorted(keys) == sorted(grads) and len(keys) == 24
    for k in keys:
        ref = g["grad/" + k]
        scale = float(np.abs(ref).max())
        assert scale > 0, k
        # same graph, same kernels: equal up to ATen's thread-dependent summation order
        np.testing.assert_allclose(grads[k], ref, rtol=0, atol=2e-5 * scale, err_msg=k)


def test_lower_bound_gradient_rule():
    """compressai's LowerBound: below the bound the gradient passes only when the step would raise x."""
    x = torch.tensor([0.5, 2.0, 0.5], requires_grad=True)
    y = O._lower_bound(x, 1.0)
    assert y.tolist() == [1.0, 2.0, 1.0]
    y.backward(torch.tensor([1.0, 1.0, -1.0]))
    assert x.grad.tolist() == [0.0, 1.0, -1.0]


def test_train_loader_crops_and_batches(tmp_path):
    from PIL import Image
    from llicti_b200.image_dl import TrainImageLoader
    for i, (h, w) in enumerate([(70, 90), (40, 100), (64, 64), (30, 30), (128, 80)]):
        Image.fromarray(O.synthetic_image(h, w, i).transpose(1, 2, 0)).save(tmp_path / f"t{i}.png")
    dl = TrainImageLoader(str(tmp_path), 64, 2, patches_per_img=1, seed=3)
    batches = list(dl)
    assert len(dl) == 3 and [tuple(b.shape) for b in batches] == [(2, 3, 64, 64), (2, 3, 64, 64), (1, 3, 64, 64)]
    assert all(b.dtype == torch.float32 and 0 <= float(b.min()) and float(b.max()) <= 1 for b in batches)
    assert all(torch.equal(torch.round(b * 255) / 255, b) for b in batches)       # uint8 / 255, what ToTensor yields
    dl4 = TrainImageLoader([str(tmp_path)], 32, 5, patches_per_img=4, seed=3)
    (b,) = list(dl4)
    assert tuple(b.shape) == (5, 4, 3, 32, 32)
    # another epoch draws other crops
    assert not torch.equal(list(dl)[0], batches[0])


def test_train_loss_list_is_differentiable():
    from llicti_b200.rate import TrainRLossList
    s = [torch.rand(2, 9, 4, 4, requires_grad=True), torch.rand(2, 9, 2, 2, requires_grad=True)]
    loss, table = TrainRLossList().forward(2 * 3 * 8 * 8, s)
    loss.backward()
    assert len(table) == 2 and len(table[0]) == 9
    assert torch.allclose(s[0].grad, torch.full_like(s[0], 3 / (2 * 3 * 8 * 8)))
    ref = sum(float(t.sum()) for t in s) / (2 * 3 * 8 * 8) * 3
    assert abs(float(loss) - ref) < 1e-5
    f, _ = TrainRLossList().forward(2 * 3 * 8 * 8, [t.detach() for t in s])
    assert isinstance(f, float) and abs(f - ref) < 1e-5
