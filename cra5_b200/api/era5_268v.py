"""Channel layout of the 268-variable CRA5 frame: 7 pressure-level variables x 37 levels followed by 9 single-level
variables (reference: cra5/api/cra5_268v_config.py:41-54). Also usable as a config file for `cra5_api(config=...)`."""

PRESSURE_VARS = ["z", "q", "u", "v", "t", "r", "w"]
SINGLE_VARS = ["v10", "u10", "v100", "u100", "t2m", "tcc", "sp", "tp", "msl"]
PRESSURE_LEVELS = [1000., 975., 950., 925., 900., 875., 850., 825., 800., 775., 750., 700., 650., 600., 550., 500.,
                   450., 400., 350., 300., 250., 225., 200., 175., 150., 125., 100., 70., 50., 30., 20., 10., 7., 5.,
                   3., 2., 1.]

# names the reference config exposes (cfg.vnames / cfg.total_levels / cfg.pressure_level)
vnames = dict(pressure=PRESSURE_VARS, single=SINGLE_VARS)
total_levels = PRESSURE_LEVELS
pressure_level = total_levels


def channel_names():
    names = [f"{v}_{int(l)}" for v in PRESSURE_VARS for l in PRESSURE_LEVELS]
    return names + list(SINGLE_VARS)
