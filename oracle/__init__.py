"""CPU oracle for the smb-vision 3D-ViT MIM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``smb_vision_b200`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
as the thing the CPU baseline times — never as a product fallback.

Parity pinning: the reference ships no golden vectors or numerical tests for
this path, so the restatement in :mod:`oracle.videomae_oracle`
is pinned against outputs of the reference itself, produced in the authoring
container by ``oracle/make_golden.py`` (which imports
``/root/reference/src/models/videomae/modeling_videomae.py`` and upstream
``transformers.VideoMAEForPreTraining``) and committed under ``tests/golden/``.
"""
