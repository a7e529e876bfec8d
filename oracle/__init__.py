"""CPU oracle of the thor-slam ingest path  -  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  Nothing
under ``thor_slam_b200/`` imports it; the product path has no CPU fallback.

What it restates (numpy / OpenCV on the CPU), stage by stage, with the
reference file:line each function follows, is listed in DESIGN.md section 3.

PARITY PIN STATUS
-----------------
* Conventions the reference implements itself - ``Extrinsics`` 4x4 direction,
  ``world_T_camera = rig_T_source @ source_T_camera``, ``CameraRig`` frame-set
  selection, stream ordering, ``BGR2RGB``/mono8 pass-through, distortion-model
  choice, ``P`` baseline term, ``RDF_TO_FLU_MATRIX``, URDF joint -> 4x4 - are
  **pinned**: ``tests/golden/*.npz|json`` hold outputs of the reference itself,
  produced in the build container by ``tests/golden/make_golden.py`` (which
  imports ``/root/reference`` with ROS / depthai stubbed out), and
  ``tests/test_oracle_golden.py`` checks the oracle against every one of them,
  including the README known answer ``rdf_to_flu @ [1,0,0,1] = [0,-1,0,1]``.
* Arithmetic the reference delegates to third parties - ``depthai``
  ``getCvFrame()`` NV12 conversion, OpenCV ``cvtColor``, cuVSLAM's undistortion,
  nvblox's back-projection - has **no golden vector in the reference**
  ("parity unpinned" by the reference's own tests).  For those stages the oracle
  *is* OpenCV 4.13 (the reference's own declared dependency,
  ``thor_slam/requirements.txt:4``) called exactly as the reference calls it,
  next to an independent pure-numpy restatement of the same integer/fp
  arithmetic; the two are checked against each other exhaustively
  (``tests/test_oracle_golden.py::test_cv_arithmetic_fixtures``) and frozen as golden fixtures.
"""
