"""Mirror of the reference's symmetry.py: the 7 listed dihedral maps (+ the unused
right_diagonal), with the transform fused into the CUDA plane gather
(csrc/rules.cu k_export_planes) and the "reverse" policy gather that re-uses the
forward map (quirk Q8: rot90/rot270 policies come back rotated 180 degrees)."""
from random import choice
import numpy as np

SYMMETRY_NAMES = ["_id", "left_diagonal", "vertical_axis", "horizontal_axis",
                  "rotation_90", "rotation_180", "rotation_270"]     # symmetry.py:117-125
RIGHT_DIAGONAL = 7                                                   # defined, unit-tested, not listed


def sym_index(size, sym):
    """Board gather table: out.flat[i] = in.flat[table[i]]."""
    S, m = size, size - 1
    y, x = np.divmod(np.arange(S * S), S)
    sy, sx = [(y, x), (x, y), (y, m - x), (m - y, x), (x, m - y), (m - y, m - x), (m - x, y), (m - x, m - y)][sym]
    return (sy * S + sx).astype(np.int64)


def random_symmetry_predict(model, board):
    """symmetry.py:127-132 for the reference's host-array protocol, computed by the
    engine's fused gather kernels."""
    from .play import _engine
    b = np.asarray(board)
    n, S = b.shape[0], b.shape[1]
    sym = choice(range(len(SYMMETRY_NAMES)))
    e = _engine(S, n)
    e.import_boards(b.astype(np.int32))
    planes = e.export_planes(0, 0, n, sym=sym).cpu().numpy()
    symm_policy, value = model.predict_on_batch(planes)
    policy = e.policy_unsym(np.asarray(symm_policy, dtype=np.float32), sym=sym).cpu().numpy()
    return policy, value
