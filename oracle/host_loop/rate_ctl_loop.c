/* ORACLE / BASELINE (test infrastructure, never on the product path).
 *
 * The reference's controller call as its host sees it (airgym/envs/base/hovering.py:246-250, ctl_mode "rate"):
 *     self.parallel_rate_control.set_q_world(root_quats_cpu.numpy().astype(np.float64))
 *     cmd_thrusts = torch.tensor(self.parallel_rate_control.update(actions_cpu.numpy().astype(np.float64),
 *                                                                  ang_vel.numpy().astype(np.float64), 0.01))
 * i.e. float64 numpy arrays handed to a C++ object that walks the envs ONE BY ONE on a single thread.  rlPx4Controller is
 * not vendored in the reference (oracle/px4_controller.py header), so this restates the same PX4 rate loop + quad-X mixer as
 * oracle/px4_controller.py:57-76,126-129 in plain C with that calling shape.  Used by bench.py's cpu_baseline to show what the
 * per-env host loop and the marshalling around it cost ("reference-shaped" baseline), and checked against the
 * vectorised torch restatement in tests/test_oracle_host_loop.py.
 *
 * state [n,6]: rate integrator (3) | previous body rate (3).  Gains as scalars (PX4 defaults are isotropic per axis triple). */
#include <stdint.h>

typedef struct RateCtlGains {
    double rate_p[3], rate_i[3], rate_d[3];
    double rate_i_fade, rate_int_lim;
} RateCtlGains;

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* q_wxyz [n,4], actions [n,4] = (wx, wy, wz set-points, collective thrust), angvel_w [n,3] world-frame rates → cmd [n,4] */
void rate_ctl_update(const RateCtlGains* g, int64_t n, const double* q_wxyz, const double* actions, const double* angvel_w,
                     double dt, double* state, double* cmd) {
    for (int64_t e = 0; e < n; ++e) {
        const double w = q_wxyz[4 * e], x = q_wxyz[4 * e + 1], y = q_wxyz[4 * e + 2], z = q_wxyz[4 * e + 3];
        /* rotation matrix of the (not re-normalised) quaternion, pytorch3d quaternion_to_matrix convention */
        const double s2 = 2.0 / (w * w + x * x + y * y + z * z);
        const double R[9] = {1 - s2 * (y * y + z * z), s2 * (x * y - z * w), s2 * (x * z + y * w),
                             s2 * (x * y + z * w), 1 - s2 * (x * x + z * z), s2 * (y * z - x * w),
                             s2 * (x * z - y * w), s2 * (y * z + x * w), 1 - s2 * (x * x + y * y)};
        const double* ww = angvel_w + 3 * e;
        double wb[3];
        for (int i = 0; i < 3; ++i) wb[i] = R[i] * ww[0] + R[3 + i] * ww[1] + R[6 + i] * ww[2]; /* R^T w */
        double* st = state + 6 * e;
        double tau[3];
        for (int i = 0; i < 3; ++i) {
            const double err = actions[4 * e + i] - wb[i];
            const double wdot = (wb[i] - st[3 + i]) / dt;
            tau[i] = g->rate_p[i] * err + st[i] - g->rate_d[i] * wdot;
            const double ef = err / g->rate_i_fade;
            double fade = 1.0 - ef * ef;
            if (fade < 0) fade = 0;
            st[i] = clampd(st[i] + fade * g->rate_i[i] * err * dt, -g->rate_int_lim, g->rate_int_lim);
            st[3 + i] = wb[i];
        }
        const double T = actions[4 * e + 3];
        cmd[4 * e + 0] = clampd(T - tau[0] - tau[1] - tau[2], 0.0, 1.0);
        cmd[4 * e + 1] = clampd(T + tau[0] + tau[1] - tau[2], 0.0, 1.0);
        cmd[4 * e + 2] = clampd(T + tau[0] - tau[1] + tau[2], 0.0, 1.0);
        cmd[4 * e + 3] = clampd(T - tau[0] + tau[1] + tau[2], 0.0, 1.0);
    }
}
