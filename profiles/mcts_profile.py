import cProfile, pstats, sys, os, torch
sys.path.insert(0, os.getcwd())
import chess_deep_q_b200 as cdq
torch.manual_seed(42)
net = cdq.ChessQNetwork().cuda()
cdq.self_play_game(net, max_moves=2, iterations=8)
pr = cProfile.Profile(); pr.enable()
cdq.self_play_games(net, n_games=32, max_moves=10, iterations=48)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
