"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference's algorithms for the multi-view aggregation hot path
(DeepViewAgg @ 41543bc: torch_points3d/modules/multimodal/pooling.py, core/multimodal/*.py,
utils/multimodal.py) plus a restatement of the documented semantics of `torch_scatter`
(un-vendored third-party dependency, version unpinned by install.sh:125 => 2.0.5-2.0.7).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker / CPU baseline. The product (deepviewagg_b200/)
never imports it.

Pinning: the reference ships NO tests or golden vectors for this path.
The oracle is pinned against (a) the four known-answer snippets the reference does hold
(pooling.py:913-921 softmax, image.py:2350-2390 CSR round trip, utils/multimodal.py:326-379
lex ops, pooling.py:870 empty->0) and (b) outputs of the reference's own modules executed in
the build container through oracle/ref_loader.py, committed as tests/golden/*.npz together
with the generating script oracle/make_golden.py.  The reference's arithmetic inside
torch_scatter itself is restated from its documentation (parity for that dependency is
"restated, not executed": torch_scatter is not installable here).
"""
