"""BASELINE config 5 / SURVEY 8f rank 2 on the GPU: the contrastive encoder's forward pass as the reference runs it --
modules left in training mode, BatchNorm on batch statistics -- and ContrastiveLoss.forward, through the frameID mirror and the
C ABI, against the reference's recorded outputs and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import kat_inputs
from oracle import contrastive as ocon

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "contrastive_kat.npz"))


def build_nets(kat):
    from frameID.net import FrameConvNet, FrameLinearNet
    conv_net = FrameConvNet(hidden_channels=32, n_conv_layers=3)
    linear_net = FrameLinearNet(n_layers=3, input_size=32, hidden_size=32, output_size=8)
    conv_net.load_state_dict({k[5:]: torch.from_numpy(kat[k]) for k in kat.files if k.startswith("conv.")}, strict=False)
    linear_net.load_state_dict({k[7:]: torch.from_numpy(kat[k]) for k in kat.files if k.startswith("linear.")}, strict=False)
    return conv_net.to("cuda"), linear_net.to("cuda")


def test_training_mode_forward_matches_reference(kat):
    """learn_contrasts.py:100-108: x = cat(x_t1, x_t2); intermediate = conv_net(x); res = linear_net(intermediate).
    Twice: on the float32 CUDA-core kernels (summation order only) and on the tensor-core path (16-bit operands, batch statistics
    taken from the stored 16-bit activations)."""
    conv_net, linear_net = build_nets(kat)
    assert conv_net.training and linear_net.training           # never put in .eval(), as in the reference script
    n = int(kat["pairs"])
    x = torch.from_numpy(np.concatenate([kat_inputs.smooth_images(n, seed=21), kat_inputs.smooth_images(n, seed=22)])).cuda()
    for tensor_cores, tol_trunk, tol_chain in ((False, 1e-4, 1e-3), (True, 2e-2, 1e-1)):
        conv_net.batchstats_tensor_cores = tensor_cores
        inter = conv_net(x)
        res = linear_net(inter)
        d_inter = float(np.abs(inter.cpu().numpy() - kat["intermediate"]).max())
        assert d_inter <= tol_trunk, (tensor_cores, d_inter)
        # chained (BatchNorm over 12 samples divides by the batch deviation of each feature, which magnifies the trunk's differences)
        d_chain = float(np.abs(res.cpu().numpy() - kat["projection"]).max())
        assert d_chain <= tol_chain, (tensor_cores, d_chain)
        print(f"contrastive forward, tensor cores {tensor_cores}: trunk {d_inter:.2e}, chained {d_chain:.2e}")
    # the projection head on the reference's own intermediate: one stage, no error carried in
    d_head = float(np.abs(linear_net(torch.from_numpy(kat["intermediate"]).cuda()).cpu().numpy() - kat["projection"]).max())
    assert d_head <= 1e-4, d_head
    # eval() switches the same modules to the running statistics (fresh modules: mean 0, var 1): a different function
    conv_net.eval()
    assert np.abs(conv_net(x).cpu().numpy() - kat["intermediate"]).max() > 1e-2


def test_contrastive_loss_matches_reference(kat):
    from frameID.metrics import ContrastiveLoss
    n = int(kat["pairs"])
    proj = torch.from_numpy(kat["projection"]).cuda()
    loss, logits_ab, labels = ContrastiveLoss(batch_size=n).to("cuda")(proj)
    assert abs(float(loss) - float(kat["loss"])) <= 1e-4
    assert np.abs(logits_ab.cpu().numpy() - kat["logits_ab"]).max() <= 1e-5
    assert labels.tolist() == list(range(n))
    loss, logits_ab, _ = ContrastiveLoss(batch_size=n, temperature=0.5, h_norm=False).to("cuda")(proj)
    assert abs(float(loss) - float(kat["loss_t05_nonorm"])) <= 2e-4
    assert np.abs(logits_ab.cpu().numpy() - kat["logits_ab_t05_nonorm"]).max() <= 1e-4 * np.abs(kat["logits_ab_t05_nonorm"]).max()
    with pytest.raises(RuntimeError):
        ContrastiveLoss(batch_size=n + 1).to("cuda")(proj)


@pytest.mark.parametrize("pairs,dim", [(1, 8), (32, 8), (33, 5), (700, 16)])
def test_contrastive_loss_against_oracle(pairs, dim):
    from cutdet import engine
    rng = np.random.default_rng(pairs * 31 + dim)
    x = rng.normal(size=(2 * pairs, dim)).astype(np.float32)
    loss, ab = engine.contrastive_loss(torch.from_numpy(x).cuda(), temperature=0.7)
    want_loss, want_ab = ocon.contrastive_loss(x, temperature=0.7)
    assert abs(float(loss) - float(want_loss)) <= 1e-4 * max(1.0, abs(float(want_loss)))
    assert np.abs(ab.cpu().numpy() - want_ab).max() <= 1e-5


def test_batchstats_full_size_batch():
    """The reference's own batch (2 x 32 frames of 144 x 256): output is normalised per feature -- mean beta, variance gamma^2."""
    from frameID.net import FrameConvNet
    torch.manual_seed(0)
    net = FrameConvNet(hidden_channels=32, n_conv_layers=3).to("cuda")
    x = torch.from_numpy(kat_inputs.smooth_images(64, seed=5)).cuda()
    w = {"conv." + k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    want = ocon.forward_batchstats(w, x.cpu().numpy(), 1)
    for tensor_cores, tol in ((False, 1e-4), (True, 2e-2)):
        net.batchstats_tensor_cores = tensor_cores
        y = net(x).cpu().numpy()            # avgpool 1x1 of the last BatchNorm'd map (5 x 9 positions)
        assert np.abs(y - want).max() <= tol, (tensor_cores, float(np.abs(y - want).max()))
        assert np.abs(y.mean(axis=0)).max() <= 1e-3        # gamma = 1, beta = 0 on a fresh module


def test_batchstats_whole_classifier_and_large_batches():
    """A glued conv + linear net (the prod architecture) left in training mode: tensor-core path for up to 148 frames, the float32
    kernels beyond (the statistics span the batch) -- both against the oracle."""
    from frameID.net import load_default_net
    from oracle import net as onet
    net, params = load_default_net()
    net.to("cuda")
    assert net.training
    w = {("conv." if i == 0 else "linear.") + k: v.detach().cpu().numpy() for i in (0, 1) for k, v in net[i].state_dict().items()}
    for batch, tol in ((40, 5e-2), (200, 1e-3)):
        x = kat_inputs.smooth_images(batch, seed=batch)
        want = ocon.forward_batchstats(w, x, params["avg_pool_size"])
        got = net(torch.from_numpy(x).cuda()).cpu().numpy()
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= tol * max(1.0, float(np.abs(want).max())), (batch, float(np.abs(got - want).max()))


def test_eval_mode_trunk_alone_runs_on_tensor_cores(kat):
    """The contrastive encoder's trunk is a BARE FrameConvNet (learn_contrasts.py:68-70: no FC layer behind it in the same module);
    after .eval() it takes the tensor-core path too, its 32 pooled features coming out of the head kernel with the avg-pool folded
    into a matrix.  Against the float32 oracle with the same running statistics, at the 16-bit-operand tolerance."""
    from oracle import net as onet
    conv_net, _ = build_nets(kat)
    with torch.no_grad():                       # non-trivial running statistics
        for m in conv_net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    conv_net.eval()
    assert conv_net._native().uses_tensor_cores(144, 256)
    x = kat_inputs.smooth_images(40, seed=9)
    got = conv_net(torch.from_numpy(x).cuda()).cpu().numpy()
    w = {"conv." + k: v.detach().cpu().numpy() for k, v in conv_net.state_dict().items()}
    want = onet.forward_f32(w, x, 1)
    assert got.shape == want.shape == (40, 32)
    assert np.abs(got - want).max() <= 0.05 * max(1.0, float(np.abs(want).max())), float(np.abs(got - want).max())
