"""CPU oracle for the ReVisionLLM segment-scoring inference path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, in plain fp32 torch/numpy on the CPU, the arithmetic the reference
runs for the hot path: the `<video>` embedding splice,
the `mm_projector` / `ClipEncoder` adapters, the Llama decoder stack the
reference reaches through `transformers` (pinned 4.41.2, not vendored in the
reference tree), greedy decode bookkeeping, entropy statistics, CLIP cosine
top-k pooling, the window builders and the stage-2 selection rule.  Every
function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import it - as the checker, never as the
thing that is measured as the product or shipped.  `revisionllm_b200/` never
imports from here.

Parity pinning: the reference ships no tests, golden vectors or fixtures for
this path, so the oracle is pinned against outputs of the
reference itself: `tests/golden/make_golden.py` imports the reference's own
modules from /root/reference (through a namespace shim, see
`oracle/ref_shim.py`) together with the installed `transformers` Llama, runs
them on seeded inputs and commits the input/output vectors under
`tests/golden/`.  `tests/test_oracle_golden.py` checks this restatement against
those vectors on every CPU run.
"""
