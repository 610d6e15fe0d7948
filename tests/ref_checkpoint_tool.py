"""Test infrastructure (CPU, this container only): drives the REFERENCE's own OptimizedChessAI.save_model /
load_model (chess_ai.py:454-485) from /root/reference with python-chess, matplotlib and tqdm stubbed -- none of
them is touched by the checkpoint code.  Used by tests/test_checkpoint.py; skipped where /root/reference is absent.

    python tests/ref_checkpoint_tool.py write <path>      # reference agent, two real train() steps, save_model
    python tests/ref_checkpoint_tool.py read <path> <npz> # reference load_model, dump what it restored
"""
import sys
from unittest.mock import MagicMock

REF = "/root/reference"


def _reference_ai():
    import importlib
    for name in ("chess", "chess.engine", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.gridspec",
                 "tqdm", "scipy", "scipy.stats", "pandas"):
        try:
            importlib.import_module(name)
        except ImportError:
            sys.modules[name] = MagicMock()
    sys.path.insert(0, REF)
    import torch
    torch.manual_seed(42)
    import chess_ai
    return chess_ai, torch


def main():
    chess_ai, torch = _reference_ai()
    import numpy as np
    ai = chess_ai.OptimizedChessAI(training_games=3)
    agent = ai.dqn_agent
    if sys.argv[1] == "write":
        rng = np.random.default_rng(0)
        for i in range(96):
            s = torch.from_numpy((rng.random((12, 8, 8)) < 0.04).astype(np.float32))
            ns = torch.from_numpy((rng.random((12, 8, 8)) < 0.04).astype(np.float32))
            agent.memory.append((s, None, float(rng.normal() * 0.01), ns, float(i % 17 == 0)))
        for _ in range(2):
            ai.loss_history.append(agent.train())
        agent.update_target_network()
        agent.train()
        agent.epsilon = 0.0875
        ai.game_history.append({"moves": ["e2e4"], "result": "1/2-1/2", "move_count": 1, "white_scores": [0.1],
                                "black_scores": [-0.1], "final_score": 0.1})
        ai.evaluation_history.append(0.1)
        ai.epsilon_history.append(0.0875)
        ai.move_count_history.append(1)
        ai.save_model(sys.argv[2])
    else:
        ai.load_model(sys.argv[2])
        out = {"epsilon": np.float64(agent.epsilon)}
        for k, v in agent.q_network.state_dict().items():
            out["q." + k] = v.cpu().numpy()
        for k, v in agent.target_q_network.state_dict().items():
            out["t." + k] = v.cpu().numpy()
        osd = agent.optimizer.state_dict()
        for i, st in osd["state"].items():
            out["m.%d" % i] = st["exp_avg"].cpu().numpy()
            out["v.%d" % i] = st["exp_avg_sq"].cpu().numpy()
            out["step.%d" % i] = np.float64(float(st["step"]))
        out["lr"] = np.float64(osd["param_groups"][0]["lr"])
        np.savez(sys.argv[3], **out)


if __name__ == "__main__":
    main()
