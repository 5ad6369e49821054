"""PI/PD/PID attitude controller used as the deterministic driver of the known-answer test.

TEST INFRASTRUCTURE.  Restates the reference's baseline controller
(magpie/libs/pyfly/pyfly/pid_controller.py:4-108: PI on airspeed -> throttle, PD on roll -> aileron, PID on pitch ->
elevator, gains and limits as there), vectorised over a batch of envs with numpy.
"""
import numpy as np


class BatchPID:
    def __init__(self, n, dt=0.01):
        self.k_p_V, self.k_i_V = 0.5, 0.1
        self.k_p_phi, self.k_i_phi, self.k_d_phi = 1.0, 0.0, 0.5
        self.k_p_theta, self.k_i_theta, self.k_d_theta = -4.0, -0.75, -0.1
        self.delta_a_min, self.delta_a_max = np.radians(-30), np.radians(30)
        self.delta_e_min, self.delta_e_max = np.radians(-30), np.radians(35)
        self.dt = dt
        self.n = n
        self.ref = np.zeros((n, 3))          # phi_r, theta_r, va_r
        self.int_va = np.zeros(n)
        self.int_roll = np.zeros(n)
        self.int_pitch = np.zeros(n)

    def reset(self, idx=None):
        idx = slice(None) if idx is None else idx
        self.int_va[idx] = 0
        self.int_roll[idx] = 0
        self.int_pitch[idx] = 0

    def set_reference(self, ref, idx=None):
        idx = slice(None) if idx is None else idx
        self.ref[idx] = ref

    def get_action(self, phi, theta, va, omega):
        """phi, theta, va: [n]; omega: [n,3].  Returns [n,3] = (elevator, aileron, throttle) commands."""
        e_V_a = va - self.ref[:, 2]
        e_phi = phi - self.ref[:, 0]
        e_theta = theta - self.ref[:, 1]
        self.int_va = self.int_va + self.dt * e_V_a
        self.int_roll = self.int_roll + self.dt * e_phi
        self.int_pitch = self.int_pitch + self.dt * e_theta
        delta_t = 0 - self.k_p_V * e_V_a - self.k_i_V * self.int_va
        delta_a = -self.k_p_phi * e_phi - self.k_i_phi * self.int_roll - self.k_d_phi * omega[:, 0]
        delta_e = 0 - self.k_p_theta * e_theta - self.k_i_theta * self.int_pitch - self.k_d_theta * omega[:, 1]
        delta_t = np.clip(delta_t, 0, 1.0)
        delta_a = np.clip(delta_a, self.delta_a_min, self.delta_a_max)
        delta_e = np.clip(delta_e, self.delta_e_min, self.delta_e_max)
        return np.stack([delta_e, delta_a, delta_t], axis=1)


# overrides applied by the reference's evaluation harness (examples/evaluate_controller.py:90-103, use_pid=True)
PID_EVAL_CONFIG_KW = {"steps_max": 1500,
                      "target": {"on_success": "done", "success_streak_fraction": 1, "success_streak_req": 100,
                                 "states": {0: {"bound": 5}, 1: {"bound": 5}, 2: {"bound": 2}}},
                      "action": {"scale_space": False}}
PID_EVAL_SIM_KW = {"turbulence": False, "turbulence_intensity": "None"}


def scenario_state21(init_state):
    """Golden scenario rows (roll..velocity_w, Va, alpha, beta, elevator, aileron, throttle, wind_n/e/d) ->
    FW_NSTATE_INJECT layout.  PyFly.reset ignores Va/alpha/beta/elevator/aileron (pyfly.py:1273-1282, :637-643);
    elevons are not in the test set and therefore start at their (0, 0) init range; throttle is taken."""
    init_state = np.asarray(init_state, dtype=np.float64)
    out = np.zeros((init_state.shape[0], 21))
    out[:, :12] = init_state[:, :12]
    out[:, 14] = init_state[:, 17]
    out[:, 18:21] = init_state[:, 18:21]
    return out
