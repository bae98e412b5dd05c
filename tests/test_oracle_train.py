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


def test_checkpoint_trained_on_the_b200_path_works_in_the_reference_model():
    """tests/golden/ckpt_A_b200_trained.npz: llicti_A trained from a fresh initialisation by THIS repo's `mode: train`
    (tools/train_b200_ckpt.py --epochs 14 on a B200: 672 steps of llicti_backward_dev + Adam, 19 s).  In the reference's own
    forward (the oracle's bit-exact restatement, CPU) its rate on the recipe's eight validation images is the 11.99 bpp
    the GPU's validate() logged at the end of that training; the checkpoint the unmodified reference trained on the same
    images (tests/golden/ckpt_A_trained.npz, a longer run) gives 11.76."""
    import os
    from conftest import GOLDEN
    from llicti_b200.synth import synthetic_image
    ocfg = O.OracleConfig()
    torch.set_num_threads(4)

    def rate(npz):
        with np.load(os.path.join(GOLDEN, npz)) as z:
            net = O.OracleNet(ocfg, {k: z[k] for k in z.files})
        tot = []
        for i in range(8):                                      # the validation set of tools/train_*_ckpt.py
            img = synthetic_image(160, 160, 5000 + 384 + 1 + i, noise=(0.7, 1.5, 2.5, 4.0)[i % 4])
            tot.append(sum(float(t.sum(dtype=np.float64)) for t in O.forward_self_informations(ocfg, net, img)) / img.size * 3)
        return float(np.mean(tot))

    assert abs(rate("ckpt_A_b200_trained.npz") - 11.990) < 0.02
    assert abs(rate("ckpt_A_trained.npz") - 11.755) < 0.02


def test_hand_wired_weights_sit_on_relu_kinks_generic_weights_do_not(monkeypatch):
    """Why gradient parity is tested with jittered weights: the stand-in weights' interpolation path (taps of exactly
    +-0.25) gives pre-activations that are mathematically ZERO wherever four neighbours cancel, so whether a ReLU passes the
    gradient there is decided by rounding noise of the inputs -- the float lifting's fl(fl(r/255) - fl(b/255)) against the
    same sample snapped to integer / 255 (an ulp apart).  With generic weights the two inputs give the same gradients."""
    ocfg = O.OracleConfig(dwtlevels=(0, 1), chs=60)
    rgb = np.stack([O.synthetic_image(24, 40, 33 + i, noise=3.0) for i in range(2)])
    torch.set_num_threads(4)
    orig = O.float_ycocg_r

    def worst(sd):
        _, g_float = O.train_loss_and_grads(ocfg, sd, rgb)
        monkeypatch.setattr(O, "float_ycocg_r", lambda x: torch.round(orig(x) * 255) / 255)
        _, g_snap = O.train_loss_and_grads(ocfg, sd, rgb)
        monkeypatch.setattr(O, "float_ycocg_r", orig)
        return max(float(np.abs(g_float[k] - g_snap[k]).max()) / (float(np.abs(g_float[k]).max()) + 1e-30) for k in g_float)

    assert worst(O.jittered_state_dict(ocfg, seed=1337)) < 1e-4
    assert worst(O.synthetic_state_dict(ocfg, seed=1337)) > 1e-3
