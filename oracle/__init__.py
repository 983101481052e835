"""CPU oracle for the Cut-Detection per-frame hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithms of the reference's hot path
(frame preprocessing -> CNN -> segmentation -> CSV) so the CUDA kernels can be
checked against them.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package (``cut-detection_b200/``) never does.

Pinning status: the reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, produced in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference/frameID`` unmodified) and committed under
``tests/golden/``.  ``tests/test_oracle_*.py`` replay those vectors.

Modules
-------
preprocess    cv2.resize(INTER_LINEAR) fixed-point restatement + BGR->RGB /255
              (reference: frameID/data.py:197-228; arithmetic lives in OpenCV).
net           FrameConvNet / FrameLinearNet forward, fp32 (torch CPU functional)
              and an independent float64 numpy restatement
              (reference: frameID/net.py:11-189).
segmentation  run-length encoding, glue_orphans, combine_adjacent_segments, CSV
              (reference: frameID/segmentation.py:12-196).
"""
