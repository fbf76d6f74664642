"""CPU oracle for the restoration-network forward/backward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the thing shipped.  The product path
(``frequency-wised_all-in-one_image_restoration_model_b200``) never imports it
and raises if its CUDA library is missing.

What it is: a plain fp32 PyTorch-on-CPU *functional restatement* of the
reference's ``net/`` modules.  Every function takes a ``state_dict``-style
mapping ``sd`` carrying the reference's own parameter names plus a key prefix,
so the very same weights drive the reference, the oracle and the CUDA path.
Each function cites the reference ``file:line`` it follows.

Parity pin: the reference publishes no golden vectors, so the
oracle is pinned against outputs of the reference itself: ``tools/make_golden.py``
imports ``/root/reference/net`` in the build container (timm stubbed, ``.cuda()``
neutralised), fills its parameters with :func:`oracle.detfill.fill_state` and
commits the resulting tensors under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them through this package.  The one
piece with *no* runnable reference is DCNv2 (``net/utils/deform_conv.py:64``
``assert False``; mmcv absent and unpinned): for it the oracle restates the
published DCNv2 definition and is cross-checked against
``torchvision.ops.deform_conv2d`` 0.26 — **parity unpinned** for that op.
"""
