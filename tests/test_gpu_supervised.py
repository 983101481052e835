"""The forward pass of the reference's supervised training / validation loop (training_scripts/supervised_training.py:134-215,
SURVEY 8f rank 4) on the native path against the same loop in plain PyTorch on the CPU:

  training step     conv_net.train(); linear_net.train(); pred = linear_net(conv_net(x)); loss = CrossEntropyLoss(sum)(pred, y)
  validation step   .eval(); pred; loss; pc = max(pred, 1)[1]; correct[c] / total[c] per class

Forward only: the optimiser step (loss.backward(), optimizer.step()) is out of scope, and with it the side effect torch's
BatchNorm has in training mode (moving running_mean / running_var towards the batch statistics): the native training-mode forward
normalises with the batch statistics and leaves the module's buffers alone."""
import numpy as np
import pytest
import torch

import kat_inputs
from oracle import net as onet
from oracle.reference_path import build_torch_net

pytestmark = pytest.mark.gpu


def _mirror_nets(weights, params):
    from frameID.net import FrameConvNet, FrameLinearNet
    conv = FrameConvNet(hidden_channels=params["conv_channels"], n_conv_layers=params["conv_layers"], average_pool_size=params["avg_pool_size"])
    lin = FrameLinearNet(n_layers=params["linear_layers"], input_size=params["conv_channels"] * params["avg_pool_size"] ** 2,
                         hidden_size=params["linear_size"], output_size=params["linear_output_size"])
    conv.load_state_dict({k[len("conv."):]: torch.from_numpy(v.copy()) for k, v in weights.items() if k.startswith("conv.")}, strict=False)
    lin.load_state_dict({k[len("linear."):]: torch.from_numpy(v.copy()) for k, v in weights.items() if k.startswith("linear.")}, strict=False)
    return conv.to("cuda"), lin.to("cuda")


def test_training_and_validation_forward(prod_weights):
    from frameID.metrics import SummedCrossEntropy
    weights, params = prod_weights
    conv, lin = _mirror_nets(weights, params)
    ref = build_torch_net(weights, params["avg_pool_size"])
    criterion_ref = torch.nn.CrossEntropyLoss(reduction="sum")
    criterion = SummedCrossEntropy()
    x = torch.from_numpy(kat_inputs.smooth_images(48, seed=5))
    rng = np.random.default_rng(0)
    y = torch.from_numpy(rng.integers(0, 3, 48))
    # training step: BatchNorm on batch statistics (float32 CUDA-core kernels where the modules run one by one)
    conv.train(); lin.train(); ref.train()
    with torch.no_grad():
        want_pred = ref(x)
        want_loss = criterion_ref(want_pred, y)
        pred = lin(conv(x.cuda()))
        loss = criterion(pred, y.cuda())
    assert float((pred.cpu() - want_pred).abs().max()) <= 5e-2            # tensor-core trunk: 16-bit operands
    assert abs(float(loss) - float(want_loss)) <= 2e-2 * max(1.0, abs(float(want_loss)))
    # the loss kernel itself, on the reference's own predictions: float32 rounding only
    assert abs(float(criterion(want_pred.cuda(), y.cuda())) - float(want_loss)) <= 1e-5 * max(1.0, abs(float(want_loss)))
    # validation step.  torch's BatchNorm moved its running statistics during the training-mode forward above; the native
    # forward does not (no optimiser step follows it here, see the module docstring), so validate against the shipped statistics
    ref = build_torch_net(weights, params["avg_pool_size"])
    conv.eval(); lin.eval(); ref.eval()
    with torch.no_grad():
        want_pred = ref(x)
        want_loss = criterion_ref(want_pred, y)
        pc = torch.max(want_pred, dim=1)[1]
        want_correct = torch.stack([torch.sum(pc[y == c] == y[y == c]) for c in range(3)])
        want_total = torch.stack([torch.sum(y == c) for c in range(3)])
        loss, correct, total = SummedCrossEntropy.accuracy_counts(want_pred.cuda(), y.cuda())
        pred = lin(conv(x.cuda()))
    assert abs(float(loss) - float(want_loss)) <= 1e-5 * max(1.0, abs(float(want_loss)))
    assert torch.equal(correct.cpu(), want_correct) and torch.equal(total.cpu(), want_total)
    assert float((pred.cpu() - want_pred).abs().max()) <= 5e-2


def test_cross_entropy_edge_cases():
    from cutdet import engine
    # large logits (the log-sum-exp must not overflow), a single class, an empty batch, out-of-range labels
    logits = torch.tensor([[1000.0, 0.0, -1000.0], [-50.0, -50.0, -50.0], [3.0, 3.0, 3.0]], device="cuda")
    y = torch.tensor([0, 1, 2], device="cuda")
    want = torch.nn.functional.cross_entropy(logits.cpu(), y.cpu(), reduction="sum")
    loss, correct, total = engine.cross_entropy_sum(logits, y, with_counts=True)
    assert abs(float(loss) - float(want)) <= 1e-5
    assert correct.cpu().tolist() == [1, 0, 0] and total.cpu().tolist() == [1, 1, 1]     # ties: first index wins
    assert float(engine.cross_entropy_sum(torch.zeros((0, 3), device="cuda"), torch.zeros(0, dtype=torch.int64, device="cuda"))) == 0.0
    with pytest.raises(IndexError):
        engine.cross_entropy_sum(logits, torch.tensor([0, 3, 1], device="cuda"))
    rng = np.random.default_rng(1)
    big = torch.from_numpy(rng.normal(0, 4, (100_003, 8)).astype(np.float32))
    yb = torch.from_numpy(rng.integers(0, 8, 100_003))
    want = torch.nn.functional.cross_entropy(big.double(), yb, reduction="sum")
    assert abs(float(engine.cross_entropy_sum(big.cuda(), yb.cuda())) - float(want)) <= 1e-5 * float(want)
