"""Which payloads and which component chains the polarimetry fixtures cover (shared by the
generator, which needs the reference, and the tests, which must not)."""

NAMES = ["simple_calcite", "incident_calcite", "azimuthal_calcite", "fullsweep_quartz", "c2_dispersion_moo3",
         "dispersion_iso_exit", "multilayer_incident"]

# (incident state, components in beam order); "sample" marks the structure
MUELLER_CHAINS = {
    "A": (("linear", {"angle": 45}), [("linear_polarizer", 20), ("anisotropic_sample",), ("quarter_wave_plate", 30)]),
    "B": (("circular", {"handedness": "right"}), [("anisotropic_sample",), ("half_wave_plate", 10), ("linear_polarizer", 75)]),
    "C": (("elliptical", {"alpha": 30, "ellipticity": 20}), [("anisotropic_sample",)]),
}
JONES_CHAIN = (("elliptical", {"alpha": 30, "ellipticity": 20}),
               [("quarter_wave_plate", 15), ("sample",), ("rotator", 12), ("linear_polarizer", 60)])
# (a_s, a_p) of the arbitrary incident state used for FieldProfile.summary
JONES_PAIR = (0.6 - 0.3j, 0.2 + 0.7j)
