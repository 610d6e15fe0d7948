"""chess-deep-q leaf evaluation, B200-native.

Drop-in for the reference's leaf-evaluation call surface (thistleknot/chess-deep-q):
board_utils.board_to_tensor, evaluation.fast_evaluate_position, neural_network.ChessQNetwork /
DQNAgent and mcts._evaluate_position, all running as hand-written sm_100a CUDA kernels behind the
C ABI in include/cdq.h.  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from . import _lib  # noqa: F401
from .board_utils import BOARD_DTYPE, Move, pack_board, pack_boards, board_to_tensor, boards_to_tensor_batch  # noqa: F401
from .evaluation import fast_evaluate_position, evaluate_batch, PIECE_VALUES  # noqa: F401
from .neural_network import ChessQNetwork, DQNAgent  # noqa: F401
from .mcts import (evaluate_position, evaluate_positions, LeafEvaluator, BatchedRussianDollMCTS, ParallelRussianDollMCTS,  # noqa: F401
                   DeviceForest,
                   self_play_game, self_play_games, search_many, move_from_records, move_uci)

__version__ = "0.1.0"
