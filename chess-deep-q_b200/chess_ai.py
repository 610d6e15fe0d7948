"""Training orchestration surface: mirrors the parts of the reference's chess_ai.py that touch the
hot path -- the checkpoint dict (save_model / load_model, chess_ai.py:454-485) and the self-play
loop that feeds the replay memory and trains once per ply (chess_ai.py:117-250, 346-413).  Plots,
the Stockfish ladder, CSV export and the interactive prompts are out of scope."""
import time

import torch

from .neural_network import DQNAgent

# keys of the reference's checkpoint dict, in its order (chess_ai.py:456-469)
CHECKPOINT_KEYS = ("q_network_state_dict", "target_network_state_dict", "optimizer_state_dict", "epsilon",
                   "game_history", "evaluation_history", "loss_history", "epsilon_history", "move_count_history",
                   "total_games_trained", "training_games", "save_timestamp")


def save_checkpoint(filename, dqn_agent, game_history=(), evaluation_history=(), loss_history=(), epsilon_history=(),
                    move_count_history=(), training_games=0):
    """torch.save of exactly the dict the reference writes (chess_ai.py:456-469): state_dicts with the
    reference's keys and shapes, the optimizer state in torch.optim.Adam's format."""
    torch.save({
        "q_network_state_dict": dqn_agent.q_network.state_dict(),
        "target_network_state_dict": dqn_agent.target_q_network.state_dict(),
        "optimizer_state_dict": dqn_agent.optimizer.state_dict(),
        "epsilon": dqn_agent.epsilon,
        "game_history": list(game_history),
        "evaluation_history": list(evaluation_history),
        "loss_history": list(loss_history),
        "epsilon_history": list(epsilon_history),
        "move_count_history": list(move_count_history),
        "total_games_trained": len(game_history),
        "training_games": training_games,
        "save_timestamp": time.time(),
    }, filename)


def load_checkpoint(filename, dqn_agent, map_location=None):
    """chess_ai.py:471-485: restores both networks, the optimizer and epsilon from a checkpoint written
    by save_checkpoint OR by the reference's own save_model; returns the whole dict."""
    ckpt = torch.load(filename, map_location=map_location or dqn_agent.device, weights_only=False)
    dqn_agent.q_network.load_state_dict(ckpt["q_network_state_dict"])
    dqn_agent.target_q_network.load_state_dict(ckpt["target_network_state_dict"])
    dqn_agent.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    dqn_agent.epsilon = ckpt["epsilon"]
    dqn_agent.q_network.invalidate_packed()
    dqn_agent.target_q_network.invalidate_packed()
    return ckpt


class OptimizedChessAI:
    """chess_ai.py:23-58 constructor and attributes; the board is a packed record with its history
    (board_utils.Game) instead of a python-chess Board."""

    def __init__(self, training_games=20, verbose=False, use_aha_learning=False):
        self.dqn_agent = DQNAgent(use_aha_learning=use_aha_learning)
        self.training_games = training_games
        self.game_history = []
        self.evaluation_history = []
        self.loss_history = []
        self.epsilon_history = []
        self.verbose = verbose
        self.move_count_history = []
        self.num_cpu_cores = self.dqn_agent.num_cpu_cores

    def save_model(self, filename="chess_ai_model.pth"):
        save_checkpoint(filename, self.dqn_agent, self.game_history, self.evaluation_history, self.loss_history,
                        self.epsilon_history, self.move_count_history, self.training_games)

    def load_model(self, filename="chess_ai_model.pth", continue_training=False):
        """Like the reference, `continue_training` is accepted and ignored (chess_ai.py:471-485)."""
        load_checkpoint(filename, self.dqn_agent)
        if self.verbose:
            print("Model loaded from %s" % filename)
