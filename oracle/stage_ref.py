"""Pack the reference modules of the hot path into oracle/_ref/reference_path.zip (git-ignored; NOT
gpurun-ignored, so the archive travels to the GPU box with the snapshot like a built .so does).

Why: the reference is pure Python, there is nothing to compile -- "building" it for the checker
means making the unmodified modules importable where /root/reference does not exist. With the
archive present `bench.py --impl reference` and the `cpu_baseline` leg run the reference's own
models/qwen3.py + layers/attention_torch_native.py (cpu_baseline.kind = "reference"); without it
they fall back to the oracle port (kind = "port"). Nothing here enters git history, and nothing
under nano-vllm-ascend_b200/ imports it.

  python -m oracle.stage_ref            # build container only; called by __graft_entry__.build()
"""
from __future__ import annotations

import hashlib
import json
import os
import zipfile

SRC = os.environ.get("B200_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
ARCHIVE = os.path.join(DST, "reference_path.zip")

# exactly the modules oracle/ref_runtime.py imports; the NPU-bound ones
# (layers/attention.py, engine/model_runner.py, engine/llm_engine.py) are deliberately absent
FILES = [
    "nanovllm/layers/attention_torch_native.py",
    "nanovllm/layers/layernorm.py",
    "nanovllm/layers/rotary_embedding.py",
    "nanovllm/layers/activation.py",
    "nanovllm/layers/sampler.py",
    "nanovllm/layers/embed_head.py",
    "nanovllm/layers/linear.py",
    "nanovllm/utils/context.py",
    "nanovllm/utils/logger.py",
    "nanovllm/models/qwen3.py",
    "nanovllm/models/qwen3_moe.py",
    "nanovllm/models/llama.py",
    "nanovllm/engine/block_manager.py",
    "nanovllm/engine/scheduler.py",
    "nanovllm/engine/sequence.py",
    "nanovllm/sampling_params.py",
    "nanovllm/config.py",
]


def stage() -> bool:
    if not os.path.isdir(os.path.join(SRC, "nanovllm")):
        return False
    manifest = {}
    os.makedirs(DST, exist_ok=True)
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        # explicit directory entries: zipimport only treats a directory as a namespace package
        # (the reference has no __init__.py below nanovllm/) when the archive lists it
        for d in sorted({os.path.dirname(rel) + "/" for rel in FILES} | {"nanovllm/"}):
            z.writestr(zipfile.ZipInfo(d, date_time=(2020, 1, 1, 0, 0, 0)), b"")
        for rel in FILES:
            src = os.path.join(SRC, rel)
            if not os.path.isfile(src):
                continue
            with open(src, "rb") as f:
                data = f.read()
            z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), data)   # unmodified bytes
            manifest[rel] = hashlib.sha256(data).hexdigest()
        z.writestr(zipfile.ZipInfo("MANIFEST.json", date_time=(2020, 1, 1, 0, 0, 0)),
                   json.dumps({"source": SRC, "sha256": manifest}, indent=1, sort_keys=True))
    return True


if __name__ == "__main__":
    print("packed" if stage() else f"{SRC} not present: nothing packed", "->", ARCHIVE)
