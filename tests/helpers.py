import numpy as np

CYCLE_FIELDS = ("decision", "mode", "borrow", "commit_rank", "ps_flavor", "ps_res_mode", "ps_tried_idx", "ps_count", "node_usage")


def assert_cycle_equal(got, want, snap=None, fields=CYCLE_FIELDS):
    for f in fields:
        g, w = getattr(got, f), getattr(want, f)
        if not np.array_equal(g, w):
            bad = np.argwhere(np.asarray(g) != np.asarray(w))[:5]
            raise AssertionError(f"{f} differs at {bad.tolist()}: got {np.asarray(g)[tuple(bad[0])]} want {np.asarray(w)[tuple(bad[0])]}")
    # targets compared per entry as ordered lists
    assert np.array_equal(got.tgt_start, want.tgt_start), "tgt_start differs"
    n = int(want.tgt_start[-1])
    assert np.array_equal(got.tgt_adm[:n], want.tgt_adm[:n]), "tgt_adm differs"
    assert np.array_equal(got.tgt_reason[:n], want.tgt_reason[:n]), "tgt_reason differs"
