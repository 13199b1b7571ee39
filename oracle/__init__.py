"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the gcm-filters hot path.

Nothing in the product package (``gcm_filters_b200``) imports this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may use it, and there only as the checker / the reported CPU
baseline, never as the thing that is shipped or measured as the product.

Contents
--------
``np_oracle``    independent numpy restatement of the reference Laplacians
                 (``gcm_filters/kernels.py``) and Chebyshev step loop
                 (``gcm_filters/filter.py:154-291``), written in explicit index form.
``fixtures``     the reference test-suite fixtures restated (``tests/conftest.py``),
                 plus the synthetic benchmark inputs of SURVEY.md section 8(d).
``zarr_golden``  decoder for the reference's blosc-lz4 zarr-v2 golden arrays.
``ref_loader``   imports the live reference from ``/root/reference`` with an xarray
                 stub.  Only usable in the build container (the tree does not exist
                 on the GPU box); used to pin the oracle and to generate
                 ``tests/golden/*.npz``.

Parity status: PINNED.  ``np_oracle`` reproduces all 18 golden arrays of the reference
test-suite (``tests/test_data_kernels``, ``tests/test_data_filter``), the two FilterSpec
known-answer tests (``tests/test_filter.py:23-79``) and is bit-identical to the live
reference on every grid type (see ``tests/test_oracle.py`` and ``tests/golden/make_golden.py``).
MOM5U / MOM5T have no golden in the reference; they are pinned only by outputs of
the live reference captured in ``tests/golden/ref_outputs.npz``.
"""
