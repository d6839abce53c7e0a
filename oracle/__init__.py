"""CPU oracle for the video_features hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a plain numpy / torch-fp32 restatement of the
reference algorithm (Kamino666/video_features @ dc9df59e), used exclusively as
the *checker* by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
package ``video_features_b200`` never imports it; the product path has no CPU
fallback and fails loudly when ``libvfeat.so`` is missing.

Pinning status (also in DESIGN.md):
  * sampler, PIL resample, transforms: pinned against the reference's own
    python (imported from /root/reference in the build container) and against
    Pillow/torchvision themselves -> fixtures in tests/golden/.
  * I3D / RAFT nets: pinned against the reference modules + vendored
    checkpoints run in the build container -> fixtures in tests/golden/.
  * CLIP image tower: the reference calls the un-vendored third-party package
    ``clip`` (openai/CLIP, unpinned) whose code and weights are absent offline.
    The restatement follows the published ``clip/model.py`` algorithm and is
    cross-checked against the independent HF ``transformers`` implementation;
    versus the reference itself it is "parity unpinned".
"""
