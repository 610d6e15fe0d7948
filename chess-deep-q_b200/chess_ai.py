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

    # ------------------------------------------------------------------ self-play and training (chess_ai.py:117-250, 346-413)
    def self_play_games(self, n_games=1, max_moves=200, seeds=None):
        """`n_games` games of self-play in lockstep on the device forest, annealed like DQNAgent._get_mcts_move
        (neural_network.py:150-163) with the reference's training progress (games played / training_games,
        chess_ai.py:133): every move is stored as a transition and the network trains once per stored transition
        (chess_ai.py:187-192); the target network is synchronised every five games (chess_ai.py:243-245).
        -> list of (result, move_count) like the reference's self_play_game."""
        from .mcts import annealed_search_parameters, self_play_games
        agent = self.dqn_agent
        progress = min(len(self.game_history) / self.training_games, 1.0) if self.training_games else 1.0
        iterations, exploration_weight, widths = annealed_search_parameters(progress)
        if seeds is None:
            base = len(self.game_history)
            seeds = [42 + base + g for g in range(n_games)]
        games = self_play_games(agent.q_network, n_games, max_moves=max_moves, iterations=iterations, samples_per_level=widths,
                                seeds=seeds, training_progress=progress, exploration_weight=exploration_weight,
                                num_workers=min(32, agent.num_cpu_cores), agent=agent)
        from .evaluation import evaluate_batch
        out = []
        for g in games:
            fl = g["result_flags"]
            white_to_move = bool(int(g["positions"][-1]["meta"]) & 1)
            if fl & 0x10:
                result = "0-1" if white_to_move else "1-0"                # the side to move is mated
            elif (fl & 0x60) or ((fl >> 24) & 6):
                result = "1/2-1/2"
            else:
                result = "*"
            final = float(evaluate_batch(g["positions"][-1:], ignore_checkmate=True)[0][0].item())    # chess_ai.py:216
            name = lambda s: "abcdefgh"[s & 7] + str((s >> 3) + 1)  # noqa: E731
            self.game_history.append({"moves": [name(f) + name(t) + ("" if p is None else " nbrq"[p]) for f, t, p in g["moves"]],
                                      "result": result, "move_count": len(g["moves"]), "white_scores": list(g["scores"]),
                                      "black_scores": [-x for x in g["scores"]], "final_score": final})
            self.loss_history.extend(g["losses"])
            self.evaluation_history.append(final)
            self.epsilon_history.append(agent.epsilon)
            self.move_count_history.append(len(g["moves"]))
            if len(self.game_history) % 5 == 0:
                agent.update_target_network()
            out.append((result, len(g["moves"])))
        return out

    def self_play_game(self, max_moves=200):
        """chess_ai.py:117-250 -> (result, move_count)"""
        return self.self_play_games(1, max_moves)[0]

    def train(self, progress_interval=1, parallel_games=1, max_moves=200):
        """chess_ai.py:346-413 without the plots and prompts: training_games games of self-play, `parallel_games` at a
        time in lockstep on the device."""
        start = time.time()
        done = 0
        while done < self.training_games:
            n = min(parallel_games, self.training_games - done)
            for result, moves in self.self_play_games(n, max_moves):
                done += 1
                if self.verbose and done % progress_interval == 0:
                    recent = self.loss_history[-50:]
                    print("Game %d/%d: %s in %d moves, avg loss %.4f, %.1fs" % (done, self.training_games, result, moves,
                                                                               sum(recent) / max(1, len(recent)), time.time() - start))
