"""GPU: host-side contracts of the engine / Trainer that the reference's users rely on --
torch optimizers see gradients, checkpoints resume bit-identically."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = {"model": {"type": "epd", "message_passing_num": 2, "hidden_size": 32, "node_input_size": 2, "output_size": 2,
                 "edge_input_size": 3},
       "index": {"feature_index_start": 0, "feature_index_end": 2, "output_index_start": 0, "output_index_end": 2,
                 "node_type_index": 2}}


def _batch(seed=0):
    from graphphysics_b200.synthetic import cylinder_flow_batch
    return cylinder_flow_batch(2, nx=20, ny=10, seed=seed)


def test_torch_optimizer_and_clip_see_the_gradients():
    """nn.Parameter.grad must be populated (views of the flat gradient buffer), so a stock
    torch optimizer / clip_grad_norm_ built on model.parameters() works as in the reference's
    configure_optimizers (lightning_module.py:494-511)."""
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    b = _batch().to(dev)
    model = EncodeProcessDecode(2, 11, 3, 2, hidden_size=32).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    x = torch.randn(b.x.shape[0], 11, device=dev)
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for it in range(2):                         # second round: zero_grad(set_to_none=True) dropped the views
        opt.zero_grad()
        out = model(Data(x=x, edge_index=b.edge_index, edge_attr=b.edge_attr))
        out.square().mean().backward()
        grads = model.engine.grads_by_name()
        for name, p in model.named_parameters():
            assert p.grad is not None, name
            assert torch.equal(p.grad, grads[name]), name
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        assert torch.isfinite(total) and float(total) > 0
        opt.step()
    after = model.state_dict()
    changed = sum(int(not torch.equal(before[k], after[k])) for k in before)
    assert changed == len(before), f"only {changed} of {len(before)} tensors were updated"
    eng = model.engine
    assert eng.is_bound()
    # a parameter whose storage was replaced un-binds the engine (every parameter is checked, not the first);
    # the `engine` property then builds a fresh one over the current values
    last = list(model.parameters())[-1]
    last.data = last.data.clone()
    assert not eng.is_bound()
    assert model.engine is not eng and model.engine.is_bound()


@pytest.mark.parametrize("graphed", [False, True])
def test_trainer_resume_is_bit_identical(graphed):
    """Trainer.state_dict / load_state_dict carry AdamW moments, the device step counter, the
    schedule position and the normalisers; 3 + 3 steps with a save/load in between equal 6 straight steps."""
    from graphphysics_b200.training.loop import Trainer
    dev = torch.device("cuda:0")
    batches = [_batch(s).to(dev) for s in range(6)]

    def make():
        tr = Trainer(copy.deepcopy(CFG), learning_rate=1e-3, num_steps=50, warmup=4, device=dev, seed=0)
        tr.enable_cuda_graph(graphed)
        return tr

    a = make()
    la = [float(a.training_step(b)) for b in batches]
    b1 = make()
    lb = [float(b1.training_step(b)) for b in batches[:3]]
    state = copy.deepcopy(b1.state_dict())
    b2 = make()
    b2.training_step(batches[5])                   # dirty every buffer first
    b2.load_state_dict(state)
    assert b2.step_index == 3 and b2.next_lr() == a.learning_rate * __import__(
        "graphphysics_b200.utils.scheduler", fromlist=["lr_factor"]).lr_factor(3, 4, 50)
    lb += [float(b2.training_step(b)) for b in batches[3:]]
    assert la == lb, (la, lb)
    sa, sb = a.model.state_dict(), b2.model.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert torch.equal(a.exp_avg, b2.exp_avg) and torch.equal(a.exp_avg_sq, b2.exp_avg_sq)


@pytest.mark.parametrize("graphed", [False, True])
def test_training_noise_injection(graphed):
    """transformations.preprocessing noise (preprocessing.py:177-238) drawn and applied on the device inside the step
    (SURVEY §8f N2): the caller's batch is left alone, a fresh draw is used every step (also under CUDA-graph replay),
    and a zero scale reproduces the noise-free run bit for bit."""
    from graphphysics_b200.training.loop import Trainer
    dev = torch.device("cuda:0")

    def run(scale, inject):
        cfg = copy.deepcopy(CFG)
        cfg["transformations"] = {"preprocessing": {"noise": scale, "noise_index_start": [0], "noise_index_end": [2]}}
        tr = Trainer(cfg, learning_rate=0.0, num_steps=100, warmup=1, device=dev, seed=0, inject_noise=inject)   # lr 0: weights stay
        if graphed:
            tr.enable_cuda_graph()
        b = _batch().to(dev)
        x0 = b.x.clone()
        losses = [float(tr.training_step(b)) for _ in range(4)]
        assert torch.equal(b.x, x0)
        return losses

    clean = run(0.02, False)
    noisy = run(0.02, True)
    assert all(np.isfinite(noisy))
    # lr = 0 and a repeated batch: without noise every step repeats (normaliser statistics aside), with noise the
    # loss moves from step to step because every step draws again
    assert len(set(noisy[1:])) == 3, noisy
    assert noisy != clean


@pytest.mark.parametrize("graphed", [False, True])
@pytest.mark.parametrize("precision", ["bf16", "tight"])
def test_trainer_on_degenerate_batches(graphed, precision):
    """A batch with no edges at all and a six-node batch: the full training step (normalisers, model, masked loss, clip,
    AdamW) and a one-frame roll-out run on both arithmetic modes, eagerly and from a captured graph; losses stay finite."""
    from graphphysics_b200.graph import Data
    from graphphysics_b200.training.loop import Trainer
    dev = torch.device("cuda:0")
    cfg = copy.deepcopy(CFG)
    cfg["model"]["precision"] = precision
    g = torch.Generator().manual_seed(1)
    for n, ei in ((9, torch.zeros((2, 0), dtype=torch.long)), (6, torch.tensor([[0, 1, 2, 3, 4, 5, 0], [1, 2, 3, 4, 5, 0, 3]]))):
        tr = Trainer(cfg, learning_rate=1e-3, num_steps=50, warmup=2, device=dev, seed=0)
        if graphed:
            tr.enable_cuda_graph()
        vel = torch.randn(n, 2, generator=g)
        x = torch.cat([vel, torch.zeros(n, 2)], 1)
        x[0, 2] = 4.0                                           # one INFLOW node: masked out of the loss
        b = Data(x=x, y=vel + 0.1, pos=torch.rand(n, 2, generator=g), edge_index=ei, edge_attr=torch.randn(ei.shape[1], 3, generator=g)).to(dev)
        losses = [float(tr.training_step(b)) for _ in range(3)]
        assert all(np.isfinite(losses)), (n, losses)
        r = tr.rollout([b])
        assert np.isfinite(r["val_1step_rmse"]) and tuple(r["predictions"][0].shape) == (n, 2)


def test_checkpointed_training_is_bit_identical_and_smaller():
    """training.enable_vram_optimizations (the reference's memory-optimised training, layers.py:24-36): the engine keeps only
    every 4th layer's inputs and re-runs segments in the backward -- gradients bit-identical, peak memory lower."""
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models import layers as L
    from graphphysics_b200.models.processors import EncodeProcessDecode
    from graphphysics_b200.synthetic import cylinder_flow_batch
    dev = torch.device("cuda:0")
    b = cylinder_flow_batch(8, seed=0).to(dev)
    torch.manual_seed(0)
    x, G_ = torch.randn(b.x.shape[0], 11, device=dev), torch.randn(b.x.shape[0], 2, device=dev)
    res = {}
    for flag in (False, True):
        L.set_memory_optimized_training(flag)
        try:
            torch.manual_seed(1)
            m = EncodeProcessDecode(10, 11, 3, 2, hidden_size=128).to(dev)
            assert m.engine.checkpoint_every == (4 if flag else 0)
        finally:
            L.set_memory_optimized_training(False)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats(dev)
        base = torch.cuda.memory_allocated(dev)
        out = m(Data(x=x, edge_index=b.edge_index, edge_attr=b.edge_attr))
        (out * G_).sum().backward()
        torch.cuda.synchronize()
        res[flag] = (out.detach().clone(), {k: v.clone() for k, v in m.engine.grads_by_name().items()}, torch.cuda.max_memory_allocated(dev) - base)
        del m, out
    assert torch.equal(res[True][0], res[False][0])
    assert all(torch.equal(res[True][1][k], res[False][1][k]) for k in res[False][1])
    assert res[True][2] < 0.7 * res[False][2], (res[True][2], res[False][2])


@pytest.mark.parametrize("graphed", [False, True])
def test_gradient_accumulation_matches_lightning_semantics(graphed):
    """Trainer(accumulate_grad_batches=2) (train.py:70, 289): each batch's loss counts 1/2, the optimizer, the clip and the
    LR schedule advance on every second call.  With the online normalisers frozen (they would otherwise move inside a
    window), feeding the same batch twice must land exactly where one plain step on it lands (g/2 + g/2 = g) -- bit for bit,
    eagerly and replayed."""
    from graphphysics_b200.training.loop import Trainer
    dev = torch.device("cuda:0")
    b0, b1 = _batch(0).to(dev), _batch(1).to(dev)
    warm = Trainer(copy.deepcopy(CFG), learning_rate=1e-3, num_steps=50, warmup=4, device=dev, seed=0)
    for b in (b0, b1):
        warm.training_step(b)
    stats = {k: v.clone() for k, v in warm.model.state_dict().items() if "_normalizer" in k}

    def make(k):
        tr = Trainer(copy.deepcopy(CFG), learning_rate=1e-3, num_steps=50, warmup=4, device=dev, seed=0, accumulate_grad_batches=k)
        tr.model.load_state_dict(stats, strict=False)
        for n in (tr.model._output_normalizer, tr.model._node_normalizer, tr.model._edge_normalizer):
            n._max_accumulations, n._host_calls = 0, 0            # frozen statistics
        tr.enable_cuda_graph(graphed)
        return tr

    plain, acc = make(1), make(2)
    before = acc.engine.flat.data.clone()
    for b in (b0, b1, b0):                       # three optimizer steps
        plain.training_step(b)
    for i, b in enumerate((b0, b0, b1, b1, b0, b0)):
        acc.training_step(b)
        if i == 0:
            assert torch.equal(acc.engine.flat.data, before) and acc.step_index == 0      # no update inside the window
    assert acc.step_index == plain.step_index == 3
    assert torch.equal(acc.engine.flat.data, plain.engine.flat.data)
    assert torch.equal(acc.exp_avg, plain.exp_avg) and torch.equal(acc.exp_avg_sq, plain.exp_avg_sq)
