"""the reference's checkpoint dict (chess_ai.py:454-485).

CPU part (runs where /root/reference exists, i.e. in the build container): a checkpoint WRITTEN BY THE
REFERENCE'S OWN save_model after real training steps loads through chess_ai.load_checkpoint, and a
checkpoint written by chess_ai.save_checkpoint loads through THE REFERENCE'S OWN load_model -- all twelve
keys, both state_dicts, torch.optim.Adam's state format, epsilon.  The state handling is pure torch, so it
needs no GPU.  GPU part: DQNAgent / OptimizedChessAI round trip and bit-identical training afterwards."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tests", "ref_checkpoint_tool.py")
HAVE_REFERENCE = os.path.exists("/root/reference/chess_ai.py")


def _cpu_agent(cdq):
    """What load_checkpoint / save_checkpoint touch of a DQNAgent, without a CUDA device."""
    from chess_deep_q_b200.train import FusedAdam
    a = types.SimpleNamespace(device=torch.device("cpu"), epsilon=0.1)
    a.q_network, a.target_q_network = cdq.ChessQNetwork(), cdq.ChessQNetwork()
    a.optimizer = FusedAdam(a.q_network, lr=1e-3)
    return a


def _tool(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, TOOL] + list(args), capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.skipif(not HAVE_REFERENCE, reason="needs the reference checkout (build container only)")
def test_reference_written_checkpoint_loads_and_ours_loads_into_the_reference(tmp_path):
    import chess_deep_q_b200 as cdq
    from chess_deep_q_b200 import chess_ai
    ref_pth, ours_pth, dump = str(tmp_path / "ref.pth"), str(tmp_path / "ours.pth"), str(tmp_path / "dump.npz")
    _tool("write", ref_pth)
    raw = torch.load(ref_pth, map_location="cpu", weights_only=False)
    assert tuple(raw.keys()) == chess_ai.CHECKPOINT_KEYS          # the twelve keys, in the reference's order
    agent = _cpu_agent(cdq)
    ckpt = chess_ai.load_checkpoint(ref_pth, agent, map_location="cpu")
    assert agent.epsilon == 0.0875 and ckpt["total_games_trained"] == 1 and len(ckpt["loss_history"]) == 2
    for k, v in raw["q_network_state_dict"].items():
        assert torch.equal(agent.q_network.state_dict()[k], v), k
    for k, v in raw["target_network_state_dict"].items():
        assert torch.equal(agent.target_q_network.state_dict()[k], v), k
    assert not torch.equal(raw["q_network_state_dict"]["fc2.bias"], raw["target_network_state_dict"]["fc2.bias"])
    osd = agent.optimizer.state_dict()
    assert agent.optimizer.step_count == 3 and set(osd["state"]) == set(raw["optimizer_state_dict"]["state"])
    for i, st in raw["optimizer_state_dict"]["state"].items():
        assert torch.equal(osd["state"][i]["exp_avg"], st["exp_avg"]) and torch.equal(osd["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(osd["state"][i]["step"]) == float(st["step"]) == 3.0
    # and back: our dict through the reference's own load_model
    agent.epsilon = 0.0625
    chess_ai.save_checkpoint(ours_pth, agent, game_history=ckpt["game_history"], loss_history=ckpt["loss_history"], training_games=3)
    assert tuple(torch.load(ours_pth, map_location="cpu", weights_only=False).keys()) == chess_ai.CHECKPOINT_KEYS
    _tool("read", ours_pth, dump)
    d = np.load(dump)
    assert float(d["epsilon"]) == 0.0625 and float(d["lr"]) == 1e-3
    for k, v in raw["q_network_state_dict"].items():
        assert np.array_equal(d["q." + k], v.numpy()), k
        assert np.array_equal(d["t." + k], raw["target_network_state_dict"][k].numpy()), k
    for i, st in raw["optimizer_state_dict"]["state"].items():
        assert np.array_equal(d["m.%d" % i], st["exp_avg"].numpy()) and np.array_equal(d["v.%d" % i], st["exp_avg_sq"].numpy())
        assert float(d["step.%d" % i]) == 3.0


def test_checkpoint_dict_round_trip_without_the_reference(tmp_path):
    """Same dict against classes re-declared from the reference's four layers and torch.optim.Adam (what its
    save_model serialises), so the format is also pinned where /root/reference is absent."""
    import chess_deep_q_b200 as cdq
    from chess_deep_q_b200 import chess_ai
    import torch.nn as nn

    class RefNet(nn.Module):                     # neural_network.py:16-27
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(12, 32, kernel_size=3, padding=1)
            self.conv2 = nn.Conv2d(32, 64, kernel_size=3, padding=1)
            self.fc1 = nn.Linear(64 * 8 * 8, 256)
            self.fc2 = nn.Linear(256, 1)

    torch.manual_seed(1)
    q, t = RefNet(), RefNet()
    opt = torch.optim.Adam(q.parameters(), lr=1e-3)
    for _ in range(2):
        opt.zero_grad()
        sum((p * p).sum() for p in q.parameters()).backward()
        opt.step()
    path = str(tmp_path / "ref_like.pth")
    torch.save({"q_network_state_dict": q.state_dict(), "target_network_state_dict": t.state_dict(),
                "optimizer_state_dict": opt.state_dict(), "epsilon": 0.09, "game_history": [], "evaluation_history": [],
                "loss_history": [], "epsilon_history": [], "move_count_history": [], "total_games_trained": 0,
                "training_games": 20, "save_timestamp": 0.0}, path)
    agent = _cpu_agent(cdq)
    chess_ai.load_checkpoint(path, agent, map_location="cpu")
    assert agent.epsilon == 0.09 and agent.optimizer.step_count == 2
    for k, v in q.state_dict().items():
        assert torch.equal(agent.q_network.state_dict()[k], v)
    out = str(tmp_path / "ours.pth")
    chess_ai.save_checkpoint(out, agent, training_games=20)
    back = torch.load(out, map_location="cpu", weights_only=False)
    opt2 = torch.optim.Adam(RefNet().parameters(), lr=5e-4)
    opt2.load_state_dict(back["optimizer_state_dict"])            # torch's own Adam accepts it
    for i, st in opt.state_dict()["state"].items():
        assert torch.equal(opt2.state_dict()["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
    assert opt2.state_dict()["param_groups"][0]["lr"] == 1e-3


@pytest.mark.gpu
def test_agent_checkpoint_round_trip_continues_training_identically(tmp_path, oracle):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import random
    import chess_deep_q_b200 as cdq
    from chess_deep_q_b200.chess_ai import OptimizedChessAI
    boards = oracle.playouts(201, seed=8)
    _, sc, fl = oracle.evaluate(boards)

    def fill(agent):
        for i in range(200):
            done = bool(fl[i + 1] & 7)
            agent.store_transition(boards[i:i + 1], None, 0.01 * np.tanh(sc[i + 1] / 10.0), None if done else boards[i + 1:i + 2], done)

    torch.manual_seed(3)
    a = OptimizedChessAI(training_games=5)
    fill(a.dqn_agent)
    random.seed(1)
    a.loss_history += [a.dqn_agent.train() for _ in range(3)]
    a.dqn_agent.update_target_network()
    a.loss_history.append(a.dqn_agent.train())
    path = str(tmp_path / "model.pth")
    a.save_model(path)
    b = OptimizedChessAI(training_games=5)
    b.load_model(path)
    fill(b.dqn_agent)
    assert b.dqn_agent.epsilon == a.dqn_agent.epsilon and b.dqn_agent.optimizer.step_count == 4
    for k, v in a.dqn_agent.q_network.state_dict().items():
        assert torch.equal(b.dqn_agent.q_network.state_dict()[k], v), k
        assert torch.equal(b.dqn_agent.target_q_network.state_dict()[k], a.dqn_agent.target_q_network.state_dict()[k]), k
    for ag in (a.dqn_agent, b.dqn_agent):              # same minibatches from here on: same seed and draw counter
        ag._device_memory()
        ag._dmem["seed"] = 1234
        ag._dmem["counter"].zero_()
        ag._graph_step = None                          # the graph holds the seed as a kernel argument
    la = [a.dqn_agent.train() for _ in range(2)]
    lb = [b.dqn_agent.train() for _ in range(2)]
    # same kernels on the same bits; fp32 atomics (loss, weight gradients) make the last digits order dependent
    assert la == pytest.approx(lb, rel=1e-5)
    assert float((a.dqn_agent.optimizer.flat - b.dqn_agent.optimizer.flat).abs().max()) <= 2.1e-3   # two Adam steps of lr 1e-3
