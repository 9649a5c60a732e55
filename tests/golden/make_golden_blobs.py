"""Golden LMDB-value blobs written by the REFERENCE's own writer (run in the container that has /root/reference):

    python tests/golden/make_golden_blobs.py

`dumps_npz` is lifted out of /root/reference/revisionllm/data/convert_h5_to_lmdb.py (videos: {"features": [T,768] fp32},
compress=True, :30-37) and /root/reference/revisionllm/data/feature_extraction/mad_clip_text_extractor.py (queries:
token_features / cls_features, :103-107) with `ast` - the scripts themselves open LMDB environments and h5 files at
import time, so they cannot be imported - and run on small seeded arrays.  The blobs and the arrays they hold go to
tests/golden/feature_blobs.npz; tests/test_host_logic.py reads them back through revisionllm_b200.features.FeatureStore."""
import ast
import io
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def lift(path, name):
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np, "io": io}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def main():
    video_writer = lift("/root/reference/revisionllm/data/convert_h5_to_lmdb.py", "dumps_npz")
    query_writer = lift("/root/reference/revisionllm/data/feature_extraction/mad_clip_text_extractor.py", "dumps_npz")
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((7, 768)).astype(np.float32)
    tok = rng.standard_normal((5, 768)).astype(np.float32)              # the extractor stores fp32 (:103-104)
    cls = rng.standard_normal((768,)).astype(np.float32)
    video_blob = video_writer({"features": feats}, compress=True)       # convert_h5_to_lmdb.py:35-36
    query_blob = query_writer({"token_features": tok, "cls_features": cls}, compress=True)
    np.savez(os.path.join(HERE, "feature_blobs.npz"), video_blob=np.frombuffer(video_blob, dtype=np.uint8),
             query_blob=np.frombuffer(query_blob, dtype=np.uint8), features=feats, token_features=tok, cls_features=cls)
    print("video blob", len(video_blob), "bytes; query blob", len(query_blob), "bytes")


if __name__ == "__main__":
    main()
