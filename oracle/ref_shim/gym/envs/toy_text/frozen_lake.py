"""Restatement of gym<=0.25 `generate_random_map` (gym itself is absent and un-pinned by the
reference's setup.py).  Pinned by the reference's cached FrozenLakeContinuous hardness files
(SURVEY.md section 4 / section 8c): the maps this produces under np.random.seed(seed) reproduce them."""
import numpy as np


def _is_valid(board, max_size):
    frontier, discovered = [(0, 0)], set()
    while frontier:
        r, c = frontier.pop()
        if (r, c) not in discovered:
            discovered.add((r, c))
            for x, y in [(1, 0), (0, 1), (-1, 0), (0, -1)]:
                r_new, c_new = r + x, c + y
                if r_new < 0 or r_new >= max_size or c_new < 0 or c_new >= max_size:
                    continue
                if board[r_new][c_new] == "G":
                    return True
                if board[r_new][c_new] != "H":
                    frontier.append((r_new, c_new))
    return False


def generate_random_map(size=8, p=0.8):
    valid = False
    board = None
    while not valid:
        p = min(1, p)
        board = np.random.choice(["F", "H"], (size, size), p=[p, 1 - p])
        board[0][0] = "S"
        board[-1][-1] = "G"
        valid = _is_valid(board, size)
    return ["".join(x) for x in board]
