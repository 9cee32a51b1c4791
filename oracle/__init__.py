"""CPU oracle for the BioMedKG GCL training-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker / the
timed CPU reference, never as the thing shipped.  ``biomedkg_b200`` never
imports this package and has no CPU fallback.

PARITY PINNING.  The reference (HySonLab/BioMedKG) ships no tests, golden
vectors or fixtures for this path, and its arithmetic
lives in three un-vendored third-party packages that are absent from this
image and cannot be installed (no network):

    torch_geometric == 2.5.3   (pyproject.toml:6)   GCNConv, dropout_edge,
                                                     mask_feature, inits.uniform
    PyGCL           == 0.1.2   (pyproject.toml:8)   DualBranchContrast, InfoNCE,
                                                     SingleBranchContrast, JSD
    torch           == 2.2.0   (Makefile:2)         Linear, SDPA, BCE-with-logits

``oracle/pyg.py`` and ``oracle/pygcl.py`` restate the published algorithms of
those two packages; parity for that third-party arithmetic is therefore
UNPINNED by the reference ("parity unpinned" - see DESIGN.md).  What *is*
pinned: ``tests/golden/make_golden.py`` imports the reference's OWN modules
(``biomedkg/model/encoder.py``, ``model/gcl.py``, ``utils/fusion.py``,
``gcl_module.py``) from /root/reference with the third-party names bound to
these restatements, runs them on seeded inputs and commits the outputs under
``tests/golden/``; ``oracle/models.py`` (which does not import the reference)
must reproduce those vectors, and the closed forms are cross-checked against
the as-written mask-materialising forms and against dense hand-computable
graphs in ``tests/test_oracle.py``.

``oracle/emu.py`` is the same GRACE step with bf16 rounding inserted exactly where
the CUDA path stores bf16 (each point switchable): it separates kernel exactness
(device vs emulation) from the cost of the storage format (emulation vs fp64) and
is pinned to ``oracle/models.py`` with every rounding point off
(``tests/test_emulation.py``).  ``pygcl.infonce_l2l_blockwise`` evaluates the
closed form in row blocks so that the full-size BASELINE configurations fit.
"""
