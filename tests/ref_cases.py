"""The reference's own rule / symmetry / MCTS unit tests (test/tests.py,
test/tree_util_tests.py), restated against an abstract `api` so the SAME cases run
on the CPU oracle (tests/test_ref_golden_cpu.py) and on the CUDA engine through the
C ABI (tests/test_gpu_ref_golden.py).  All at SIZE=9, KOMI=5.5 like tests.py:4-6.

api: game_init() -> (board[1,9,9,17] int32, 1); make_play(x,y,board,color=None);
     legal_moves(board) -> int[82]; get_winner(board)."""
import numpy as np

S = 9


def _seq(api, moves):
    board, _ = api.game_init()
    for m in moves:
        api.make_play(*m[:2], board, *(m[2:]))
    return board


def case_self_suicide(api):                       # tests.py:250-267
    b = _seq(api, [(0, 0), (1, 0), (8, 9), (2, 1), (8, 8), (3, 0), (2, 0)])
    assert b[0][0][1][0] == 1 and b[0][0][1][1] == 0
    assert b[0][0][2][0] == 0 and b[0][0][2][1] == 0


def case_legal_not_suicide(api):                  # tests.py:269-282
    b = _seq(api, [(0, 0), (1, 0), (1, 1), (2, 1), (8, 8), (3, 0)])
    assert api.legal_moves(b)[2] == 0


def case_legal_suicide(api):                      # tests.py:284-297
    b = _seq(api, [(0, 1), (1, 0), (1, 1), (2, 1), (8, 8), (3, 0)])
    assert api.legal_moves(b)[2] == 1


def case_legal_suicide2(api):                     # tests.py:299-312
    b = _seq(api, [(3, 0), (1, 0), (1, 1), (2, 1), (3, 1, -1), (4, 0, -1)])
    assert api.legal_moves(b)[2] == 1


def case_legal_suicide3(api):                     # tests.py:314-330
    b = _seq(api, [(1, 2), (2, 0), (3, 1), (3, 0), (1, 1, -1), (4, 1, -1), (2, 2, -1), (3, 2, -1)])
    assert api.legal_moves(b)[10] == 1


def case_ko(api):                                 # tests.py:332-353
    b = _seq(api, [(0, 0), (1, 0), (1, 1), (2, 1), (8, 8), (3, 0), (2, 0)])
    mask = api.legal_moves(b)
    assert b[0][0][1][0] == 0 and b[0][0][1][1] == 0 and b[0][0][1][2] == 1 and b[0][0][1][3] == 0
    assert mask[1] == 1


def case_not_ko(api):                             # tests.py:355-381
    b = _seq(api, [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (8, 8), (3, 0)])
    mask = api.legal_moves(b)
    for x in (1, 2):
        assert b[0][0][x][0] == 0 and b[0][0][x][1] == 0 and b[0][0][x][2] == 1 and b[0][0][x][3] == 0
    assert mask[1] == 0 and mask[2] == 0


def case_full_board_capture(api):                 # tests.py:383-435
    board, _ = api.game_init()
    for i in range(S * S - 2):
        api.make_play(i % S, i // S, board)
        api.make_play(0, S, board)
    api.make_play(0, S, board)
    api.make_play(S - 1, S - 1, board)
    for i in range(S * S - 2):
        assert board[0][i // S][i % S][0] == 1 and board[0][i // S][i % S][1] == 0
    assert board[0][S - 1][S - 1][0] == 0 and board[0][S - 1][S - 1][1] == 1
    assert board[0][S - 1][S - 2][0] == 0 and board[0][S - 1][S - 2][1] == 0
    api.make_play(S - 2, S - 1, board)
    for i in range(S * S - 1):
        assert board[0][i // S][i % S][0] == 0 and board[0][i // S][i % S][1] == 1
    assert board[0][S - 1][S - 1][0] == 0 and board[0][S - 1][S - 1][1] == 0
    api.make_play(S - 1, S - 1, board)
    for i in range(S * S - 1):
        assert board[0][i // S][i % S][0] == 0 and board[0][i // S][i % S][1] == 0
    assert board[0][S - 1][S - 1][0] == 0 and board[0][S - 1][S - 1][1] == 1


def case_bug(api):                                # tests.py:437-481
    board, _ = api.game_init()
    blacks = [(5, 6), (6, 6), (6, 8), (7, 8), (8, 8)]
    for i in range(S * S):
        x, y = i % S, i // S
        if (x, y) in blacks:
            api.make_play(x, y, board)
            api.make_play(0, S, board)
        elif (x, y) == (6, 7):
            api.make_play(0, S, board)
            api.make_play(0, S, board)
        else:
            api.make_play(0, S, board)
            api.make_play(x, y, board)
    api.make_play(0, S, board)
    api.make_play(6, 7, board)
    for i in range(S * S - 1):
        x, y = i % S, i // S
        if (x, y) in blacks:
            assert board[0][y][x][0] == 0 and board[0][y][x][1] == 0
        else:
            assert board[0][y][x][0] == 0 and board[0][y][x][1] == 1


def _abs_board_to_tensor(api, real):
    """Embed an absolute-colour array (tests.py colour tests) in a 9x9 board, black to move."""
    board, _ = api.game_init()
    real = np.asarray(real)
    for y in range(real.shape[0]):
        for x in range(real.shape[1]):
            if real[y, x] == 1:
                board[0, y, x, 0] = 1
            elif real[y, x] == -1:
                board[0, y, x, 1] = 1
    return board


BIG = [[0, 0, 0, 1, 0, -1, 0, 0, 0], [0, 0, 0, 1, 0, -1, 0, 0, 0], [0, 0, 0, 1, 0, -1, 0, 0, 0],
       [0, 0, 0, 1, -1, 0, 0, -1, 0], [1, 1, 1, -1, 0, -1, -1, 0, 0], [0, 0, 0, 1, -1, 0, 0, -1, -1],
       [0, 0, 0, 1, 0, -1, 0, 0, 0], [0, 0, 0, 1, 0, -1, 0, 1, 0], [0, 0, 0, 0, 0, -1, 0, 0, 0]]


def case_get_winner_points(api):                  # tests.py:119-135: {0:29, 1:12, 2:11, -1:15, -2:14}
    b = _abs_board_to_tensor(api, BIG)
    w, black, white = api.get_winner(b)
    assert black == 12 + 11 and white == 15 + 14 + 5.5 and w == -1


RULE_CASES = [case_self_suicide, case_legal_not_suicide, case_legal_suicide, case_legal_suicide2, case_legal_suicide3,
              case_ko, case_not_ko, case_full_board_capture, case_bug, case_get_winner_points]
