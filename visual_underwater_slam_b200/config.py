"""The two compile-time switches of gtsam that change the arithmetic of the batch.py:337 path.

gtsam decides both in CMake, so a Python user of the reference gets whatever the installed wheel was built with
(README.md:18 pins no version).  The defaults here are the `pip install gtsam` 4.1 / 4.2 wheel configuration:

  GTSAM_TANGENT_PREINTEGRATION = ON          PreintegratedImuMeasurements integrates in the tangent space
                                             (TangentPreintegration.cpp); OFF = ManifoldPreintegration (Forster et al.)
  GTSAM_SLOW_BUT_CORRECT_BETWEENFACTOR = OFF BetweenFactor Jacobians are those of Between() only (H1 = -Ad(hx^-1), H2 = I);
                                             ON multiplies both by the derivative of Local() (dLog)

    import visual_underwater_slam_b200 as gtsam
    gtsam.set_gtsam_build(tangent_preintegration=False)      # reproduce a Manifold build

The setting is read when a PreintegratedImuMeasurements is constructed / a graph is packed, is carried in the packed
problem (`prob["options"]`) and handed to the library with vus_set_gtsam_build (include/vus.h).
"""
_BUILD = dict(tangent_preintegration=True, slow_but_correct_betweenfactor=False)


def gtsam_build():
    return dict(_BUILD)


def set_gtsam_build(tangent_preintegration=None, slow_but_correct_betweenfactor=None):
    if tangent_preintegration is not None:
        _BUILD["tangent_preintegration"] = bool(tangent_preintegration)
    if slow_but_correct_betweenfactor is not None:
        _BUILD["slow_but_correct_betweenfactor"] = bool(slow_but_correct_betweenfactor)
    return gtsam_build()
