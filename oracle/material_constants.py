"""Material constants for the ORACLE (test infrastructure only).

Restates the numbers of the reference's ``hyperbolic_optics/material_params.json``
(reference data file; consumed at ``materials.py:297-326, 520-538, 655-676``) as
a flat Python table.  The product package carries its own independent copy
(``hyperbolic_optics_b200/material_db.py``); ``tests/test_oracle_pin.py`` checks
both against the reference when ``/root/reference`` is present.

Units: wavenumbers (cm^-1).  ``tolo`` entries are
``(eps_inf, omega_TO[], gamma_TO[], omega_LO[], gamma_LO[])`` per optical axis.
"""

UNIAXIAL = {
    # name: (default f_min, default f_max, ordinary, extraordinary)
    "Quartz": (
        410.0, 600.0,
        (2.356, [393.5, 450.0, 695.0, 797.0, 1065.0, 1158.0], [2.1, 4.5, 13.0, 6.9, 7.2, 9.3],
         [403.0, 507.0, 697.6, 810.0, 1226.0, 1155.0], [2.8, 3.5, 13.0, 6.9, 12.5, 9.3]),
        (2.383, [363.5, 487.5, 777.0, 1071.0], [4.8, 4.0, 6.7, 6.8],
         [386.7, 550.0, 790.0, 1229.0], [7.0, 3.2, 6.7, 12.0]),
    ),
    "Sapphire": (
        210.0, 1000.0,
        (3.077, [384.99, 439.1, 569.0, 633.63], [3.3, 3.1, 4.7, 5.0],
         [387.6, 481.68, 629.5, 906.6], [3.1, 1.9, 5.9, 14.7]),
        (3.072, [397.52, 582.41], [5.3, 3.0], [510.87, 881.1], [1.1, 15.4]),
    ),
    "Calcite": (  # registry name "Calcite" is the *upper* variant (materials.py:1073)
        1300.0, 1600.0,
        (2.7, [712, 1410.0, 297.0, 223.0, 102.0], [5.0, 10.0, 14.4, 11.4, 5.7],
         [715, 1550.0, 381.0, 239.0, 123.0], [5.0, 10.0, 14.4, 11.4, 5.7]),
        (2.4, [871.0, 303.0, 92.0], [3.0, 9.1, 5.6], [890.0, 387.0, 136.0], [3.0, 9.1, 5.6]),
    ),
    "AlN": (
        500.0, 1050.0,
        (4.68, [669.0], [5.0], [912.0], [5.0]),
        (4.68, [608.0], [5.0], [889.0], [5.0]),
    ),
    "SiC": (
        500.0, 1050.0,
        (6.56, [797.5], [4.76], [972.3], [4.76]),
        (6.56, [797.5], [4.76], [972.3], [4.76]),
    ),
    "hBN": (
        700.0, 1700.0,
        (4.87, [1370.0], [5.0], [1610.0], [5.0]),
        (2.95, [780.0], [4.0], [830.0], [4.0]),
    ),
    "GaN": (
        450.0, 800.0,
        (5.35, [560.0], [4.0], [743.0], [4.0]),
        (5.35, [533.0], [4.0], [735.0], [4.0]),
    ),
}
# CalciteLower shares Calcite's oscillators with a different default window.
UNIAXIAL["CalciteLower"] = (860.0, 920.0) + UNIAXIAL["Calcite"][2:]

BIAXIAL = {
    # name: (f_min, f_max, x, y, z)
    "MoO3": (
        500.0, 1050.0,
        (5.78, [820.0], [4.0], [972.0], [4.0]),
        (6.07, [958.0], [2.0], [1004.0], [2.0]),
        (2.47, [545.0], [4.0], [851.0], [4.0]),
    ),
}

MONOCLINIC = {
    "GalliumOxide": {
        "range": (350.0, 800.0),
        "Au": {
            "eps_inf": 3.71,
            "amplitude": [544.9, 727.1, 592.1, 78],
            "omega_tn": [663.17, 448.66, 296.63, 154.84],
            "gamma_tn": [3.2, 10.5, 14.9, 2.4],
        },
        "Bu": {
            "eps_inf": {"xx": 3.75, "yy": 3.21, "xy": -0.08},
            "amplitude": [266.2, 406.5, 821.9, 795.7, 365.8, 164.2, 485.7, 520.7],
            "omega_tn": [743.48, 692.44, 572.52, 432.57, 356.79, 279.15, 262.34, 213.79],
            "gamma_tn": [11.0, 6.55, 12.36, 10.13, 3.83, 1.98, 1.75, 1.9],
            "alpha_tn": [47.8, 5.1, 106.0, 21.0, 144.0, 4.0, 158.5, 80.9],
        },
    },
}

N_DEFAULT_FREQ = 410  # materials.py:35
