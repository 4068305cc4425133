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


def engine_rows(blocks, meta, p64, A):
    """Canonical rows from a downloaded engine arena (sejonggo_b200.engine.Engine.download_tree)."""
    rows = []
    if not meta['valid']:
        return np.zeros(0, dtype=ROW_DTYPE)

    def rec(b, depth, f64):
        blk = blocks[b]
        for slot in range(A):
            if not (blk['exist'][slot >> 5] >> (slot & 31)) & 1:
                continue
            n = int(blk['n'][slot])
            w = np.float32(blk['w'][slot])
            mean = np.float32(w / np.float32(n)) if n > 0 else np.float32(0)
            p = float(p64[slot]) if f64 else float(blk['prior'][slot])
            busy = 2 if (blk['busy'][slot >> 5] >> (slot & 31)) & 1 else 0
            c = int(blk['child'][slot])
            rows.append((depth, slot, n, w, mean, p, busy, 1 if c >= 0 else 0))
            if c >= 0:
                rec(c, depth + 1, False)

    import sys
    sys.setrecursionlimit(10000)
    rec(0, 0, bool(meta['root_f64']))
    a = np.zeros(len(rows), dtype=ROW_DTYPE)
    for i, r in enumerate(rows):
        a[i] = r
    return a


def oracle_rows_fast(node):
    """oracle_rows through the bulk children() export (one ctypes call per expanded node): the same canonical
    rows, fast enough for 19x19 trees with thousands of expanded nodes."""
    from oracle import oracle as o
    L = o.lib()
    chunks = []

    def rec(ptr, depth):
        n = o.Node(ptr, owner=False)
        ch = n.children()
        k = len(ch['moves'])
        a = np.zeros(k, dtype=ROW_DTYPE)
        a['depth'] = depth
        a['move'], a['count'], a['value'], a['mean'] = ch['moves'], ch['counts'], ch['values'], ch['means']
        a['p'], a['busy'], a['expanded'] = ch['ps'], ch['busy'], ch['expanded']
        for i in range(k):
            chunks.append(a[i:i + 1])
            if ch['expanded'][i]:
                rec(L.orc_node_child_at(ptr, i), depth + 1)

    import sys
    sys.setrecursionlimit(10000)
    rec(node.ptr, 0)
    return np.concatenate(chunks) if chunks else np.zeros(0, dtype=ROW_DTYPE)
