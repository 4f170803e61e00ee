"""
gordo_b200 -- B200-native (sm_100a) implementation of equinor/gordo's per-machine autoencoder
anomaly path, behind gordo's own estimator surface.

    gordo_b200.machine.model.models.KerasAutoEncoder / KerasLSTMAutoEncoder / KerasLSTMForecast
    gordo_b200.machine.model.anomaly.diff.DiffBasedAnomalyDetector
    gordo_b200.builder.FleetModelBuilder
    gordo_b200.fleet.FFFleet             (thousands of Machines per launch)

All arithmetic runs in libgordo_b200.so (include/gordo_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
