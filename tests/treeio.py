"""Shared helpers: canonical tree rows (same layout as oracle/gen_golden.py)."""
import numpy as np

ROW_DTYPE = [('depth', 'i4'), ('move', 'i4'), ('count', 'i8'), ('value', 'f4'),
             ('mean', 'f4'), ('p', 'f8'), ('busy', 'i4'), ('expanded', 'i4')]


def oracle_rows(node):
    rows = []

    def rec(d, depth):
        for m, c in d.items():
            rows.append((depth, m, c['count'], np.float32(c['value']), np.float32(c['mean_value']), c['p'],
                         c['virtual_loss'], 1 if c['subtree'] else 0))
            rec(c['subtree'], depth + 1)

    rec(node.to_dict(), 0)
    a = np.zeros(len(rows), dtype=ROW_DTYPE)
    for i, r in enumerate(rows):
        a[i] = r
    return a


def rows_equal(a, b):
    if len(a) != len(b):
        return False, "row count %d != %d" % (len(a), len(b))
    for f in ('depth', 'move', 'count', 'busy', 'expanded'):
        if not np.array_equal(a[f], b[f]):
            i = int(np.nonzero(a[f] != b[f])[0][0])
            return False, "field %s differs at row %d: %r vs %r" % (f, i, a[i], b[i])
    for f in ('value', 'mean', 'p'):
        if not np.array_equal(a[f].view(np.uint32 if a[f].dtype == np.float32 else np.uint64),
                              b[f].view(np.uint32 if b[f].dtype == np.float32 else np.uint64)):
            i = int(np.nonzero(a[f] != b[f])[0][0])
            return False, "field %s differs (bitwise) at row %d: %r vs %r" % (f, i, a[i], b[i])
    return True, ""
