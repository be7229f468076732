"""CPU oracle for the local-3D-attention / VQ denoiser hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``world_modelz_b200`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` do, and there only as the checker / CPU baseline.

The oracle restates, in plain PyTorch-on-CPU / numpy, the arithmetic of the
reference modules

* ``vq-video-diffusion/local_3d_attention.py`` (``Local3dAttention``,
  ``Local3dAttentionTransformer``),
* ``vq-video-diffusion/vq.py`` (``VectorQuantizerEMA``),
* ``vq-video-diffusion/main.py:25-36`` (``VqVideoDiffusionModel``) and the
  corruption / sampling loops in ``main.py:50-117,246-259``.

Parity pinning: the reference ships no golden vectors, so the
oracle is pinned against outputs of the reference modules themselves, imported
from ``/root/reference`` in the build container by ``tests/golden/make_golden.py``
and committed as fixtures under ``tests/golden/`` (``tests/test_oracle_golden.py``
checks the oracle against them on every CPU run).
"""
