"""Parity cases shared by oracle/make_golden.py and tests/: (config file, env overrides, simulator overrides, envs,
steps, action scale).  They cover BASELINE.json configs 1-5 at oracle-sized N plus the edge paths the reference
exercises (failure, success done/new, resample, history > 1, integrator, potential reward)."""

CASES = {
    # configs[0]/[1]: default config, turbulence off
    "default": dict(config="fixed_wing_config.json", config_kw=None, sim_kw={"turbulence": False}, n=6, steps=40, amp=1.2),
    # configs[2]: turbulence + observation noise
    "turb_noise": dict(config="fixed_wing_config.json", config_kw={"observation": {"noise": {"mean": 0, "var": 0.1}}},
                       sim_kw={"turbulence": True, "turbulence_intensity": "moderate"}, n=6, steps=40, amp=1.0),
    # configs[3]: examples config (relative targets, no noise key, 12 obs)
    "examples": dict(config="fixed_wing_config_examples.json", config_kw=None,
                     sim_kw={"turbulence": True, "turbulence_intensity": "severe"}, n=4, steps=30, amp=1.0),
    # configs[4]: dev config + history/integrator/resample overrides
    "dev_history": dict(config="fixed_wing_config_dev.json",
                        config_kw={"integration_window": 10, "steps_max": 45,
                                   "observation": {"length": 5, "step": 1, "shape": "matrix",
                                                   "states": {6: {"value": "integrator"}, 7: {"value": "relative"}}},
                                   "target": {"resample_every": 20},
                                   "reward": {"factors": {1: {"type": "int_error"}}}},
                        sim_kw={"turbulence": False}, n=4, steps=100, amp=0.5),
    "norm_step2": dict(config="fixed_wing_config.json",
                      config_kw={"observation": {"length": 3, "step": 2, "normalize": True}, "steps_max": 30},
                      sim_kw={"turbulence": False}, n=4, steps=70, amp=1.3),
    # success streak: loose bounds so the streak fires; on_success new resamples, done terminates
    "success_new": dict(config="fixed_wing_config.json",
                        config_kw={"target": {"success_streak_req": 5, "success_streak_fraction": 0.6, "on_success": "new",
                                              "states": {0: {"bound": 120}, 1: {"bound": 60}, 2: {"bound": 20}}},
                                   "reward": {"form": "potential"}},
                        sim_kw={"turbulence": False}, n=4, steps=40, amp=1.0),
    "success_done": dict(config="fixed_wing_config_examples.json",
                         config_kw={"steps_max": 60, "target": {"success_streak_req": 8, "success_streak_fraction": 1,
                                                                "on_success": "done",
                                                                "states": {0: {"bound": 150}, 1: {"bound": 80}, 2: {"bound": 25}}},
                                    "reward": {"terms": {0: {"function_class": "exponential", "weight": 0.7}},
                                               "factors": {i: {"function_class": "exponential"} for i in range(5)}}},
                         sim_kw={"turbulence": False}, n=4, steps=40, amp=1.0),
    # constraint failure path: tight omega constraints make PyFly raise inside the integration
    "failure": dict(config="fixed_wing_config.json",
                    config_kw={"simulator": {"states": {6: {"constraint_min": -70, "constraint_max": 70},
                                                        7: {"constraint_min": -70, "constraint_max": 70},
                                                        8: {"constraint_min": -70, "constraint_max": 70}}}},
                    sim_kw={"turbulence": False}, n=6, steps=60, amp=1.5),
    # FwSpecGeneric paths no shipped configuration reaches: steady wind (rotated into the body frame in every RHS) and
    # the polynomial drag model, plus a constraint on a variable the shipped instantiation does not check (velocity_u)
    "wind": dict(config="fixed_wing_config.json", config_kw=None,
                 sim_kw={"turbulence": True, "turbulence_intensity": "light", "wind_magnitude_min": 2,
                         "wind_magnitude_max": 6}, n=4, steps=30, amp=1.0),
    "poly_drag": dict(config="fixed_wing_config_constraints.json", config_kw=None,
                      sim_kw={"turbulence": False, "drag_model": "polynomial"}, n=6, steps=40, amp=1.2),
    # per-episode simulator-parameter randomisation (gaussian with clips; several resets)
    "param_rand": dict(config="fixed_wing_config_randomised.json", config_kw={"steps_max": 18},
                       sim_kw={"turbulence": False}, n=6, steps=60, amp=1.0),
    "param_rand_uniform": dict(config="fixed_wing_config_randomised.json",
                               config_kw={"steps_max": 25, "simulator": {"model": {"distribution": "uniform"}}},
                               sim_kw={"turbulence": True, "turbulence_intensity": "light"}, n=4, steps=40, amp=1.0),
    # the CNN controller's configuration (examples/models/cnn_controller): 5-row matrix observation of 12 variables,
    # action space bounded by the actuator limits (low / high null), bounds_outside_cost
    "cnn": dict(config="fixed_wing_config_cnn.json", config_kw={"steps_max": 40},
                sim_kw={"turbulence": True, "turbulence_intensity": "moderate"}, n=4, steps=60, amp=1.3),
    # ---- reward / target branches no shipped configuration reaches (fixed_wing_config_zoo.json, oracle/make_configs.py) ----
    # every reward class (action value / delta / bound, state value / error / int_error, success timesteps / constant,
    # step, goal per_state / all), three function classes, potential form with shaping memory, success "new"
    "reward_zoo": dict(config="fixed_wing_config_zoo.json",
                       config_kw={"reward": {"form": "potential"},
                                  "target": {"success_streak_req": 5, "success_streak_fraction": 0.6, "on_success": "new",
                                             "states": {0: {"bound": 120}, 1: {"bound": 60}, 2: {"bound": 20}}}},
                       sim_kw={"turbulence": False}, n=6, steps=50, amp=1.2),
    # absolute form; a tight pitch bound makes goal per_state differ from goal all
    "reward_zoo_abs": dict(config="fixed_wing_config_zoo.json",
                           config_kw={"steps_max": 25,
                                      "target": {"success_streak_req": 6, "success_streak_fraction": 0.5,
                                                 "states": {0: {"bound": 100}, 1: {"bound": 4}, 2: {"bound": 20}}}},
                           sim_kw={"turbulence": True, "turbulence_intensity": "light"}, n=4, steps=40, amp=1.0),
    # linear targets (fixed_wing.py:495-500,977-978): fast roll slopes cross +-pi (wrap, :988-989); pitch linear under
    # the Va compensate class (:944-946); resampled mid-episode and at resets
    "target_linear": dict(config="fixed_wing_config_zoo.json",
                          config_kw={"steps_max": 70,
                                     "target": {"resample_every": 45,
                                                "states": {0: {"class": "linear", "slope_low": 500, "slope_high": 800},
                                                           1: {"class": "linear"}}}},
                          sim_kw={"turbulence": False}, n=6, steps=90, amp=0.8),
    # sinusoidal targets (:501-507,979-980) incl. sinusoidal pitch under Va compensate (bias branch, :947-948);
    # resampling at steps_count > 0 exercises the bias formula
    "target_sinusoidal": dict(config="fixed_wing_config_zoo.json",
                              config_kw={"steps_max": 60,
                                         "target": {"resample_every": 25,
                                                    "states": {0: {"class": "sinusoidal"}, 1: {"class": "sinusoidal"}}}},
                              sim_kw={"turbulence": False}, n=6, steps=80, amp=0.8),
    # sinusoidal Va, linear pitch, constant roll, 3-row observation (target / error rings carry the moving targets)
    "target_mixed": dict(config="fixed_wing_config_zoo.json",
                         config_kw={"steps_max": 35, "observation": {"length": 3, "step": 1},
                                    "target": {"states": {1: {"class": "linear"},
                                                          2: {"class": "sinusoidal", "amplitude_low": 1.0,
                                                              "amplitude_high": 3.0}}}},
                         sim_kw={"turbulence": True, "turbulence_intensity": "moderate"}, n=4, steps=50, amp=0.8),
    # reward.randomize_scaling (fixed_wing.py:330-334): factors with scaling = [low, high] redraw it at every reset
    "reward_rand_scaling": dict(config="fixed_wing_config_zoo.json",
                                config_kw={"steps_max": 20,
                                           "reward": {"randomize_scaling": True,
                                                      "factors": {0: {"scaling": [2.0, 5.0]}, 3: {"scaling": [40, 80]},
                                                                  6: {"scaling": [5.0, 20.0]}}}},
                                sim_kw={"turbulence": False}, n=4, steps=50, amp=1.0),
}
