/*
 * cdq.h -- C ABI of the B200-native batched leaf evaluator (libcdq.so).
 *
 * The reference (thistleknot/chess-deep-q) has no FFI layer: its boundary for this path is a
 * Python call surface.  Each entry point below names the reference call it replaces
 * (file:line relative to the reference checkout).  Python binds these with ctypes
 * (see INTEGRATION.md); nothing here mentions torch.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller; h_* is a HOST pointer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = CDQ_E_* below;
 *   - every device-pointer entry point works on the CURRENT device of the calling thread and keeps no state
 *     between calls except per-device caches (SM count, kernel attributes), so calls from a thread pool
 *     (the reference's callers, mcts.py:77) and on several devices of one process are safe.  Two exceptions,
 *     both serialised internally: cdq_train_fwd_bwd shares one side stream and its events per device (the
 *     ENQUEUE of concurrent training calls on one device is mutually exclusive; the work itself overlaps only
 *     as far as that side stream allows), and a CdqHostCtx runs one cdq_leaf_eval_host_ctx at a time;
 *   - run-time switches are read from the environment ONCE, at first use, and changed afterwards only through
 *     cdq_set_option;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef CDQ_H
#define CDQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDQ_VERSION 100 /* 0.1.0 */

/* ---- error codes (negative; positive values are cudaError_t) ---- */
#define CDQ_E_BADARG (-1)   /* NULL pointer, negative count, misaligned buffer        */
#define CDQ_E_NODEVICE (-2) /* no CUDA device / wrong architecture (needs sm_100)      */
#define CDQ_E_WORKSPACE (-3)/* workspace smaller than cdq_workspace_bytes(n)           */
#define CDQ_E_BADBOARD (-4) /* a record has >1 king per colour or overlapping boards   */
#define CDQ_E_TIMEOUT (-5)  /* a data-parallel peer never reached cdq_dp_adam_step       */

/* ---- PackedBoard: the wire format (72 bytes, little endian) -------------------------
 * Everything evaluation.py / board_to_tensor / ChessQNetwork read from a python-chess
 * Board, as the Board's own public integer attributes:
 *   bb[0..5] = board.pawns, .knights, .bishops, .rooks, .queens, .kings
 *   bb[6]    = board.occupied_co[chess.WHITE]     bb[7] = board.occupied_co[chess.BLACK]
 * bit i of a bitboard = square i, a1 = 0 ... h8 = 63 (python-chess numbering, so plane
 * `ch` of board_to_tensor (board_utils.py:33-38) is bit-for-bit bb[type] & bb[6 + colour]).
 * meta:
 *   bit  0       turn, 1 = white to move (chess.WHITE == True)
 *   bits 1..4    clean castling rights: 1 = white h1 (K), 2 = white a1 (Q),
 *                                        3 = black h8 (k), 4 = black a8 (q)
 *   bits 8..14   raw board.ep_square + 1 (0 = None)
 *   bit  16      rep3: host-computed board.is_repetition(3) (history dependent;
 *                evaluation.py:55, mcts.py:209).  0 for history-less boards.
 *   other bits   reserved, must be 0
 */
typedef struct CdqBoard {
    uint64_t bb[8];
    uint64_t meta;
} CdqBoard;

#define CDQ_META_TURN_WHITE 0x1ull
#define CDQ_META_CASTLE_WK 0x2ull
#define CDQ_META_CASTLE_WQ 0x4ull
#define CDQ_META_CASTLE_BK 0x8ull
#define CDQ_META_CASTLE_BQ 0x10ull
#define CDQ_META_EP_SHIFT 8
#define CDQ_META_EP_MASK 0x7full
#define CDQ_META_REP3 0x10000ull

/* ---- integer evaluation terms: int32 [n][CDQ_NTERMS], the unit of the bit-exact check ----
 * Row/col names follow evaluation.py:66-209. */
enum {
    CDQ_T_W_MAT = 0,   CDQ_T_B_MAT = 1,     /* material sums, evaluation.py:67-70           */
    CDQ_T_W_MOB = 2,   CDQ_T_B_MOB = 3,     /* legal-move counts, turn forced, :78-84       */
    CDQ_T_W_ATT = 4,   CDQ_T_B_ATT = 5,     /* |attacked squares|, :90-94                   */
    CDQ_T_W_KATT = 6,  CDQ_T_B_KATT = 7,    /* "king attackers" (enemy piece count), :106-120 */
    CDQ_T_W_CASTLED = 8, CDQ_T_B_CASTLED = 9,   /* king on g1/c1, g8/c8, :126-127            */
    CDQ_T_W_KCENTRE = 10, CDQ_T_B_KCENTRE = 11, /* king on d4/e4/d5/e5, :128-129             */
    CDQ_T_W_DBL = 12,  CDQ_T_B_DBL = 13,    /* doubled pawns, :149-150                      */
    CDQ_T_W_ISO = 14,  CDQ_T_B_ISO = 15,    /* isolated pawns, :156-176                     */
    CDQ_T_W_CENTRE = 16, CDQ_T_B_CENTRE = 17,   /* attacked centre squares, :185-186         */
    CDQ_T_W_EXT = 18,  CDQ_T_B_EXT = 19,    /* attacked extended centre, :187-188           */
    CDQ_T_W_DEF = 20,  CDQ_T_B_DEF = 21,    /* defended own pieces, :196-202                */
    CDQ_T_CHECK = 22,                       /* +1 black to move in check, -1 white, :207-209 */
    CDQ_T_TERMINAL = 23,                    /* CDQ_TERM_* below, :50-57 / mcts.py:202-210   */
    CDQ_NTERMS = 24
};
enum { CDQ_TERM_NONE = 0, CDQ_TERM_CHECKMATE = 1, CDQ_TERM_STALEMATE = 2,
       CDQ_TERM_INSUFFICIENT = 3, CDQ_TERM_REP3 = 4 };

/* flags word per position */
#define CDQ_F_TERM_MASK 0x7u        /* CDQ_TERM_* (first true of checkmate, stalemate, insufficient, rep3) */
#define CDQ_F_CHECK 0x8u            /* side to move is in check                                          */
#define CDQ_F_CHECKMATE 0x10u
#define CDQ_F_STALEMATE 0x20u
#define CDQ_F_INSUFFICIENT 0x40u
#define CDQ_F_REP3 0x80u
#define CDQ_F_STM_MOVES_SHIFT 8     /* bits 8..15: legal moves of the side to move (saturates at 255)     */
#define CDQ_F_WHITE_TO_MOVE 0x10000u /* copy of meta bit 0, so consumers of flags need not re-read boards */
#define CDQ_F_BADBOARD 0x80000000u  /* record violates the format (see CDQ_E_BADBOARD)                    */

/* ---- ChessQNetwork parameters (neural_network.py:16-27), fp32, PyTorch layouts ---- */
typedef struct CdqWeights {
    const float* conv1_w; /* [32,12,3,3] */
    const float* conv1_b; /* [32]        */
    const float* conv2_w; /* [64,32,3,3] */
    const float* conv2_b; /* [64]        */
    const float* fc1_w;   /* [256,4096]  column index = c*64 + row*8 + col (neural_network.py:33) */
    const float* fc1_b;   /* [256]       */
    const float* fc2_w;   /* [1,256]     */
    const float* fc2_b;   /* [1]         */
} CdqWeights;
#define CDQ_NPARAMS 1071073

/* Opaque handle: weights repacked for the tensor-core kernels (fp16, UMMA tile order).
 * Rebuild after every optimizer step / load_state_dict. */
typedef struct CdqNet CdqNet;

int cdq_version(void);
const char* cdq_error_string(int code);

/* Run-time switches, initialised once from the environment variable in brackets:
 *   "pipe"         [CDQ_PIPE, 0]          1: batches >= 512 K positions run conv + fc as one kernel (act2 stays in L2)
 *   "fc_split"     [CDQ_FC_SPLIT, 1]      0: small batches use the persistent fc kernel instead of the cluster split-K one
 *   "host_ramp"    [CDQ_HOST_RAMP, 1]     0: cdq_leaf_eval_host starts with full chunks
 *   "dp_timeout_s" [CDQ_DP_TIMEOUT_S, 600] wall-clock limit of cdq_dp_adam_step's wait for its peers
 *   "train_split_optimizer" [CDQ_TRAIN_SPLIT_OPT, 0] 1: cdq_train_step applies fc1.weight's optimizer launch under the convolution backward
 * Unknown names return CDQ_E_BADARG. */
int cdq_set_option(const char* name, int value);
int cdq_get_option(const char* name, int* value);

/* Replaces evaluation.fast_evaluate_position (evaluation.py:30-229) for a batch.
 * Any of d_terms / d_score / d_flags may be NULL.  ignore_checkmate as evaluation.py:30. */
int cdq_eval_terms(const CdqBoard* d_boards, int64_t n, int ignore_checkmate,
                   int32_t* d_terms /*[n][24]*/, double* d_score /*[n]*/,
                   uint32_t* d_flags /*[n]*/, void* stream);

/* Replaces find_threatened_squares / find_guarded_squares / calculate_attack_range /
 * calculate_movement_range (evaluation.py:368-428) for a batch: d_maps[i] = { squares attacked by
 * white, attacked by black, legal destination squares of white, of black (turn forced) }. */
int cdq_attack_maps(const CdqBoard* d_boards, int64_t n, uint64_t* d_maps /*[n][4]*/, void* stream);

/* Replaces board_utils.board_to_tensor (board_utils.py:12-40) /
 * archive boards_to_tensor_batch (archive/board_utils.py:2-10): fp32 planes [n,12,8,8]. */
int cdq_encode_planes(const CdqBoard* d_boards, int64_t n, float* d_planes, void* stream);

/* Network handle: repack fp32 PyTorch-layout parameters (device pointers) for the kernels. */
int cdq_net_create(CdqNet** out);
int cdq_net_load(CdqNet* net, const CdqWeights* d_weights, void* stream);
void cdq_net_destroy(CdqNet* net);

/* Workspace (device bytes) needed by cdq_value_fwd / cdq_leaf_eval for n positions. */
size_t cdq_workspace_bytes(int64_t n);

/* Replaces ChessQNetwork.forward(board_to_tensor(b)) (neural_network.py:29-37, mcts.py:216-219)
 * for a batch of packed boards: d_value[i] = tanh(fc2(...)), fp32. */
int cdq_value_fwd(const CdqNet* net, const CdqBoard* d_boards, int64_t n, float* d_value,
                  void* d_workspace, size_t workspace_bytes, void* stream);

/* Replaces ParallelRussianDollMCTS._evaluate_position (mcts.py:196-232) plus the handcrafted
 * score for a batch: d_leaf[i] = -1/+1 checkmate, 0 draw, else clamp(net value, -1, 1);
 * d_score / d_flags as cdq_eval_terms.  d_value_raw (optional) = unclamped network output.
 * net == NULL is the reference's `neural_net is None` branch (mcts.py:226-229): behind the same terminal
 * rules d_leaf[i] = tanh(fast_evaluate_position(board) / 10), computed in fp64 and rounded to fp32 (CUDA's
 * tanh against libm's: compared to 1e-6, not bit for bit); no workspace is needed, d_value_raw must be NULL. */
int cdq_leaf_eval(const CdqNet* net, const CdqBoard* d_boards, int64_t n,
                  float* d_leaf, double* d_score, uint32_t* d_flags, float* d_value_raw,
                  void* d_workspace, size_t workspace_bytes, void* stream);

/* Host-buffer path used by the Python drop-ins and bench `e2e`: copies h_boards to the device (pinned or
 * pageable) in double-buffered chunks on two streams, runs cdq_leaf_eval, copies results back, synchronises.
 * A CdqHostCtx owns the streams and device buffers (about 2.5 GB) on the device that was current at creation;
 * one call at a time per context, any number of contexts.  cdq_leaf_eval_host uses one lazily created context
 * per device (callers sharing it are serialised); cdq_host_release frees those. */
typedef struct CdqHostCtx CdqHostCtx;
int cdq_host_ctx_create(CdqHostCtx** out);
void cdq_host_ctx_destroy(CdqHostCtx* ctx);
int cdq_leaf_eval_host_ctx(CdqHostCtx* ctx, const CdqNet* net, const CdqBoard* h_boards, int64_t n,
                           float* h_leaf, double* h_score, uint32_t* h_flags);
int cdq_leaf_eval_host(const CdqNet* net, const CdqBoard* h_boards, int64_t n,
                       float* h_leaf, double* h_score, uint32_t* h_flags);
void cdq_host_release(void);

/* ---- DQN replay training step (replaces DQNAgent.train, neural_network.py:218-252) ----
 * cdq_train_fwd_bwd: q = net(states), q' = target_net(next_states), y = r + (1-done)*gamma*q',
 * loss = mean((q-y)^2) (loss_kind 0, the reference's F.mse_loss) or mean Huber(delta=1) (loss_kind 1),
 * backward through the whole network.  d_weights = the fp32 master parameters (PyTorch layouts) that
 * `net` was packed from; d_grads = where to ACCUMULATE the gradients (same layouts; zero them first);
 * *d_loss is accumulated too (zero it first).  A terminal next_state is the all-zero board
 * (neural_network.py:215).  mean_over: the loss and its gradients are sums over the n samples of THIS call divided
 * by mean_over; <= 0 means n (F.mse_loss's mean, the single-GPU case).  In an N-rank data-parallel step over a
 * global batch of B samples every rank passes mean_over = B / N, whatever its own shard size, so that the
 * 1/N average of the per-rank results (grad_scale = 1/N in cdq_adam_step, or cdq_dp_adam_step's own) is the
 * mean over the global batch also when B does not divide evenly.  OR loss_kind with CDQ_TRAIN_WS_PERSISTENT when d_workspace
 * is the same buffer as in the previous call with the same n AND was zero-filled before its first
 * use: the call then skips re-zeroing the (never written) borders of its board-shaped buffers.
 * The target network's forward and fc1's weight gradient run on an internal side stream, forked
 * from and joined to `stream` with events (graph-capture safe).
 * cdq_adam_step: torch.optim.Adam defaults (neural_network.py:65) over one flat fp32 buffer. */
#define CDQ_LOSS_MSE 0
#define CDQ_LOSS_HUBER 1
/* loss_kind 2: the backward pass of ChessQNetwork.forward alone, for torch.autograd (the reference calls
 * loss.backward() on whatever it builds from the network's output, neural_network.py:244-246).  d_reward holds
 * d loss / d q per sample, normalised by the caller so that max |.| <= 1 (gradients travel between kernels as
 * fp16); target_net, d_next_states and d_done are ignored (may be NULL); *d_loss accumulates sum(q * dq). */
#define CDQ_LOSS_UPSTREAM 2
#define CDQ_TRAIN_WS_PERSISTENT 0x100
#define CDQ_TRAIN_REPACK 0x200 /* re-pack `net` from d_weights inside the call (next to the target forward) */
#define CDQ_TRAIN_ZERO_LOSS 0x400 /* *d_loss is zeroed inside the call (no separate fill launch in a captured step) */
/* Replay memory on the device (replaces `random.sample(self.memory, batch_size)` + `torch.stack(...).to(device)` of
 * DQNAgent.train, neural_network.py:224-229).  The ring lives in device memory as four arrays of `capacity` entries (state
 * records, next-state records, rewards, done flags) plus d_ring = 4 x int32 {start slot, length, capacity, scratch (zero it once)}.
 * cdq_replay_append copies `count` new transitions from HOST memory into slots [first_slot, first_slot + count) (a wrapped
 * range is two calls) and publishes the ring's new geometry.  cdq_replay_sample draws `batch` DISTINCT entries uniformly
 * among the `len` logical entries and copies them into the step's input buffers: entry k of the minibatch is logical index
 * cdq_replay_permute(k, len, seed, *d_counter) -- a keyed pseudo-random permutation (four-round Feistel network with cycle
 * walking), so every thread computes its own index and the host can reproduce the draw -- and *d_counter advances by one per
 * call (both read on the device: the call is CUDA-graph friendly).  batch <= len is the caller's business: with batch > len
 * only the first len entries of the minibatch are written. */
int cdq_replay_append(CdqBoard* d_mem_states, CdqBoard* d_mem_next, float* d_mem_reward, float* d_mem_done, int32_t* d_ring,
                      const CdqBoard* h_states, const CdqBoard* h_next, const float* h_reward, const float* h_done, int64_t first_slot,
                      int64_t count, int32_t start, int32_t len, int32_t capacity, void* stream);
int cdq_replay_sample(const CdqBoard* d_mem_states, const CdqBoard* d_mem_next, const float* d_mem_reward, const float* d_mem_done,
                      const int32_t* d_ring, uint64_t seed, uint32_t* d_counter, int64_t batch, CdqBoard* d_states,
                      CdqBoard* d_next_states, float* d_reward, float* d_done, void* stream);
uint32_t cdq_replay_permute(uint32_t k, uint32_t len, uint64_t seed, uint32_t counter);
size_t cdq_train_workspace_bytes(int64_t n);
int cdq_train_fwd_bwd(const CdqNet* net, const CdqNet* target_net, const CdqWeights* d_weights,
                      const CdqBoard* d_states, const CdqBoard* d_next_states, const float* d_reward,
                      const float* d_done, int64_t n, double mean_over, float gamma, int loss_kind,
                      const CdqWeights* d_grads, float* d_loss, void* d_workspace, size_t workspace_bytes, void* stream);
int cdq_adam_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n_params, int step,
                  float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* ---- data-parallel optimizer step: reduce-scatter + Adam + all-gather in ONE kernel -------------
 * Replaces `ncclAllReduce(grads); optimizer.step()` of a data-parallel DQNAgent.train (the
 * reference trains on one device only, neural_network.py:218-252).  Every rank maps
 * every rank's flat parameter buffer, flat gradient buffer and flag row (CUDA IPC, cdq_ipc_*) and
 * passes the mapped pointers in rank order.  Rank r sums ITS slice of all gradient buffers (rank
 * order, so all replicas get identical bits), scales by 1/world, updates its slice of d_m / d_v
 * (optimizer state is sharded: only the owned slice of those arrays is meaningful), writes the new
 * parameters into every rank's parameter buffer; after the closing handshake every rank zeroes its own gradient buffer.
 * *d_step (device int32, Adam's t) is incremented by the call; d_grid_bar = 4 zero-initialised
 * device uint32 that belong to this group and are never reset ([0] arrivals, [1] launch generation,
 * [2] sticky time-out flag, [3] generation of the last complete arrival; cdq_train_step uses words 4..7 the same way for
 * its second optimizer launch, so allocate 8).  n_params_padded = parameter count
 * rounded up to a multiple of 4 (padding elements stay zero).  world = 1 is the single-GPU fused
 * Adam step with a device-side step counter (CUDA-graph friendly).  All ranks must call it once
 * per step and reach it within option "dp_timeout_s" of wall clock (default 600 s) of each other: a rank
 * whose wait runs out skips its update, sets d_grid_bar[2] and returns normally -- nothing traps, the CUDA
 * context stays usable -- and cdq_dp_adam_status (synchronises `stream`) reports CDQ_E_TIMEOUT; parameters
 * and optimizer state are then undefined and must be reloaded from a checkpoint. */
#define CDQ_MAX_RANKS 8
typedef struct CdqDpPeers {
    float* params[CDQ_MAX_RANKS];
    float* grads[CDQ_MAX_RANKS];
    uint32_t* flags[CDQ_MAX_RANKS];   /* flags[r] = rank r's row of 2 x CDQ_MAX_RANKS uint32, zero-initialised */
    /* Optional NVLink-multicast (NVLS) addresses of the SAME two buffers, i.e. one virtual address that the NVSwitch maps onto
     * every rank's copy (cuMulticast* objects; torch.distributed._symmetric_memory provides them): with both non-NULL the kernel
     * sums a gradient chunk of all ranks with ONE multimem.ld_reduce (the addition happens in the switch) and writes a parameter
     * chunk to all ranks with ONE multimem.st, instead of `world` peer loads and `world` peer stores.  NULL = unicast peer access. */
    float* mc_params;
    float* mc_grads;
} CdqDpPeers;
int cdq_dp_adam_step(const CdqDpPeers* peers, int rank, int world, float* d_m, float* d_v, int64_t n_params_padded,
                     int32_t* d_step, uint32_t* d_grid_bar, float lr, float beta1, float beta2, float eps, void* stream);
int cdq_dp_adam_status(const uint32_t* d_grid_bar, void* stream);

/* The whole replay step as ONE call: cdq_train_fwd_bwd + cdq_dp_adam_step + the repack of `net` for the next step, scheduled
 * as one dependency graph so that the optimizer and repack of fc1.weight (98 % of the parameters) run under the convolution
 * backward instead of after it: fc1's weight gradient -> [reduce-scatter +] Adam on fc1.weight [+ all-gather] -> fc1 tiles of
 * `net` on the library's side stream, the remaining 22 497 parameters after conv1's gradient.  Requirements: d_weights and
 * d_grads are the eight views, in the reference's parameter order, of the flat buffers opt->peers.params[rank] / .grads[rank]
 * (fc1.weight at floats [21 984, 1 070 560)); `net` is packed from the current parameters on entry (cdq_net_load once; every
 * call leaves it packed from the updated ones; CDQ_TRAIN_REPACK is implied and ignored).  d_grid_bar is 8 words here.
 * That schedule is option "train_split_optimizer" = 1; with 0 (default) the call is cdq_train_fwd_bwd (repack at its head) followed
 * by one cdq_dp_adam_step.  Either way `net` must not be used between the call and the completion of `stream`. */
typedef struct CdqOptimizer {
    CdqDpPeers peers;
    int32_t rank, world;
    float* d_m;
    float* d_v;
    int64_t n_params_padded;
    int32_t* d_step;
    uint32_t* d_grid_bar;
    float lr, beta1, beta2, eps;
} CdqOptimizer;
int cdq_train_step(const CdqNet* net, const CdqNet* target_net, const CdqWeights* d_weights, const CdqBoard* d_states,
                   const CdqBoard* d_next_states, const float* d_reward, const float* d_done, int64_t n, double mean_over,
                   float gamma, int loss_kind, const CdqWeights* d_grads, float* d_loss, void* d_workspace,
                   size_t workspace_bytes, const CdqOptimizer* opt, void* stream);
/* CUDA IPC plumbing for the above: allocate a zeroed device region and export its 64-byte handle;
 * map a peer's handle; unmap; free. */
int cdq_ipc_alloc(size_t bytes, void** d_ptr, void* handle64);
int cdq_ipc_open(const void* handle64, void** d_ptr);
int cdq_ipc_close(void* d_ptr);
int cdq_ipc_free(void* d_ptr);

/* Synthetic workload: n random playouts from the start position on the device.
 * Position i plays plies_i = U{0..max_plies} uniformly random legal moves (counter-based RNG
 * keyed by (seed, i, ply)); stops early at no-legal-move / insufficient material. */
int cdq_random_playouts(CdqBoard* d_boards, int64_t n, uint64_t seed, int max_plies, void* stream);

/* All legal successor positions of each board's side to move (list(board.legal_moves) + push,
 * mcts.py:166-194): d_children is [n][max_children] records, d_counts[i] = legal move count. */
int cdq_expand(const CdqBoard* d_boards, int64_t n, int max_children, CdqBoard* d_children,
               int32_t* d_counts, void* stream);

/* Same, plus the tactical sampling weight (10..1) of every child's move as categorize_moves assigns
 * it (evaluation.py:232-339); select_weighted_moves (evaluation.py:342-366) samples with them. */
int cdq_expand_weighted(const CdqBoard* d_boards, int64_t n, int max_children, CdqBoard* d_children,
                        int32_t* d_counts, uint8_t* d_weights /*[n][max_children]*/, void* stream);
/* Weighted child sampling on the device (replaces select_weighted_moves, evaluation.py:342-366):
 * for every node, `width` successive draws without replacement from its children, each with
 * probability proportional to the child's tactical weight (the output of cdq_expand_weighted);
 * nodes with at most `width` children keep all of them in order.  Draw k of node i uses
 * z = mix64(d_seeds[i] + (k+1) * 0x9E3779B97F4A7C15) (splitmix64 finalizer), r = (z * W) >> 64 with
 * W = sum of the remaining weights, and takes the first remaining child in index order whose
 * running weight sum exceeds r: integer arithmetic only, reproducible on the host.
 * d_selected is [n][width] records, d_sel_counts[i] = min(count_i, width).  max_children <= 224 (a chess position has at most
 * 218 legal moves). */
int cdq_sample_children(const CdqBoard* d_children, const int32_t* d_counts, const uint8_t* d_weights, int64_t n,
                        int max_children, int width, const uint64_t* d_seeds, CdqBoard* d_selected,
                        int32_t* d_sel_counts, void* stream);


/* ---- device-resident Russian Doll MCTS (replaces ParallelRussianDollMCTS.search / _parallel_simulation / _expand_node,
 * mcts.py:69-194, for any number of concurrent games) --------------------------------------------------
 * A CdqForest holds, in device memory, `n_games` games (position, halfmove clock, ply, transposition keys since the last
 * irreversible move) and one search tree per game.  cdq_forest_search runs `iterations` simulations per tree in waves of
 * `workers` (the reference's thread count, mcts.py:20,77): per wave one descent kernel (per level: node expansion with categorize_moves /
 * select_weighted_moves, evaluation.py:232-366; UCB selection with the annealed exploration weight, mcts.py:34-67;
 * Board.push; is_game_over / is_repetition bookkeeping, mcts.py:151,209), ONE cdq_leaf_eval over the leaves of
 * all trees (net == NULL: the heuristic branch) and one backup kernel (mcts.py:158-164).  Nothing is read back: the host
 * synchronises once per search.  Depth limit = len(samples_per_level), children per node = samples_per_level[0]
 * (mcts.py:127,188).  cdq_forest_advance plays every tree's most visited root move in its game (chess_ai.py:151).
 * All random draws are counter-based hashes of (game seed, ply, node key / wave, worker, depth): a tree's result does not
 * depend on the other trees in the batch.  Limits: workers <= 32, depth <= 8, width <= 32, 256 history keys per game. */
typedef struct CdqForest CdqForest;
typedef struct CdqForestConfig {
    int32_t n_games;         /* concurrent games = trees                                        */
    int32_t workers;         /* simulations per wave (ParallelRussianDollMCTS.num_workers)       */
    int32_t depth;           /* len(samples_per_level)                                           */
    int32_t width;           /* samples_per_level[0]                                             */
    int32_t max_iterations;  /* sizes the node pool: iterations * depth + 1 nodes per tree       */
} CdqForestConfig;
typedef struct CdqForestView {          /* device pointers into the forest (read-only for the caller) */
    int32_t n_games, workers, depth, width, max_nodes;
    const CdqBoard* game_boards;        /* [n_games] current positions (rep3 bit maintained)                  */
    const int32_t* game_halfmove_clock; /* [n_games]                                                          */
    const int32_t* game_ply;            /* [n_games]                                                          */
    const int32_t* game_active;         /* [n_games]                                                          */
    const uint32_t* game_status;        /* [n_games] after cdq_forest_advance: 1 moved, 2 fivefold, 4 75-move  */
    const int32_t* game_history_len;    /* [n_games]                                                          */
    const uint32_t* best_move;          /* [n_games] from | to << 6 | (promotion type) << 12 | special << 16; 0xffffffff none */
    const int32_t* best_visits;         /* [n_games] visit count of that root child                            */
    const int32_t* root_children;       /* [n_games] number of (sampled) root children                         */
    const uint32_t* edge_move;          /* [n_games][max_nodes][width]; node 0 of a tree is its root           */
    const int32_t* edge_visits;         /* N(state, action)                                                   */
    const double* edge_value;           /* Q(state, action), running mean                                     */
    const int32_t* node_children;       /* [n_games][max_nodes]                                               */
    const int32_t* node_count;          /* [n_games] expanded nodes                                           */
    const CdqBoard* leaf_boards;        /* [n_games][workers] leaves of the last wave                         */
    const float* leaf_values;           /* [n_games][workers]                                                 */
    const int32_t* leaf_depth;          /* [n_games][workers]                                                 */
    const int64_t* positions_evaluated; /* [n_games] ParallelRussianDollMCTS.total_positions_evaluated        */
} CdqForestView;
int cdq_forest_create(const CdqForestConfig* cfg, CdqForest** out);
void cdq_forest_destroy(CdqForest* forest);
/* Install the games: d_boards [n_games]; optional halfmove clocks, plies (ply = the reference's current_move), history
 * (d_history [n_games][history_stride] records of the positions BEFORE the current one since the last irreversible move,
 * oldest first, d_history_len [n_games]) and active flags; d_seeds [n_games] uint64. */
int cdq_forest_set_games(CdqForest* forest, const CdqBoard* d_boards, const int32_t* d_halfmove_clock, const int32_t* d_ply,
                         const CdqBoard* d_history, const int32_t* d_history_len, int history_stride, const uint64_t* d_seeds,
                         const int32_t* d_active, void* stream);
int cdq_forest_set_active(CdqForest* forest, const int32_t* d_active, void* stream);
int cdq_forest_search(CdqForest* forest, const CdqNet* net, int iterations, float exploration_weight, float training_progress,
                      int max_moves, void* stream);
int cdq_forest_advance(CdqForest* forest, void* stream);
int cdq_forest_view(const CdqForest* forest, CdqForestView* out);

/* Kernel launch counter (all cdq_* kernels launched by this process), for bench gpu_launches. */
uint64_t cdq_launch_count(void);

/* Per-kernel device timing (CUDA events on the launching stream), used by bench.py's roofline
 * line.  Kinds: 0 eval, 1 conv stack, 2 fc head, 3 encode planes, 4 playouts/expand,
 * 5 weight pack, 6 train. */
void cdq_profile_enable(int on);
int cdq_profile_collect(double* ms_per_kind /*[8]*/, uint64_t* launches_per_kind /*[8]*/);
/* same, training kernels one by one: kinds 8 head (loss/tanh/fc2), 9 fc1 weight gradient, 10 fc1 data gradient,
 * 11 conv2 backward, 12 conv1 weight gradient, 13 optimizer (kind 6 stays empty) */
int cdq_profile_collect_ex(double* ms_per_kind /*[16]*/, uint64_t* launches_per_kind /*[16]*/);

/* Parity/debug hook: runs only the conv stack; dumps ReLU(conv1) as fp16 [n][64][32] (optional)
 * and the fc1 operand tiles (fp16, ceil(n/128) MiB). */
int cdq_debug_conv(const CdqNet* net, const CdqBoard* d_boards, int64_t n, void* d_act1,
                   void* d_act2_tiles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDQ_H */
