"""Port of correct_boundary_finite_element.py onto the GPU path: the least-squares correction of the reaction
boundary (sigma, gamma) so that the finite-element total fluxes of the production problem
(advection_diffusion_mixed.solve_problem_instance, DG2 x RT3 x CG3) match the required reaction rates
4 pi 2 sigma c/(c + gamma) (:11-12) for a set of concentrations.

Same functions and call structure (required_rate :11, jacobian :15-25, finite_element_fluxes :59-86, the
least_squares call :115-118); the FreeFem++ / gmsh / dolfin-convert meshing of create_mesh (:28-56) is replaced by
the boundary-fitted structured mesh MappedRectangleMesh(n, n, r2_max, r1_max, sigma, gamma, grading=2) (grid lines
clustered at r2 = 0 and at the reaction boundary, where the exp(-4/3 pi c r2^3) layer sits), and
solve_problem_instance is closestmolecule_b200.mixed.solve_problem_instance_mixed (same signature and return).
usage: python scripts/run_boundary_correction.py [n=32] [max_nfev=6]"""
import os
import sys
import time

import numpy as np
from scipy.optimize import least_squares

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from closestmolecule_b200 import MappedRectangleMesh, structured_boundary_markers  # noqa: E402
from closestmolecule_b200.mixed import solve_problem_instance_mixed as solve_problem_instance  # noqa: E402


GRADING = 2.0


def required_rate(conc, sigma, gamma):
    return 4 * np.pi * 2 * sigma * (conc / (conc + gamma))


def jacobian(parameters, sigma_orig, gamma_orig, c_vals, r1_max, r2_max, n, t_final, dt, D1, D2, steady_state):
    """First-order model sigma*c/(c+gamma) of the flux: d/dsigma = c/(c+gamma), d/dgamma = -sigma c/(c+gamma)^2."""
    sigma, gamma = parameters
    return np.transpose(np.array([c_vals / (c_vals + gamma), -sigma * c_vals / np.square(c_vals + gamma)]))


def finite_element_fluxes(parameters, sigma_orig, gamma_orig, c_vals, r1_max, r2_max, n, t_final, dt, D1, D2,
                          steady_state):
    sigma, gamma = float(parameters[0]), float(parameters[1])
    print(f'New sigma: {sigma} new gamma: {gamma}')
    V1 = 4 * np.pi * np.power(r1_max - sigma, 3) / 3
    mesh = MappedRectangleMesh(n, n, r2_max, r1_max, sigma, gamma, grading=GRADING)
    bdry = structured_boundary_markers(mesh)
    finite_ele_fluxes = []
    for c in c_vals:
        print(f'Current concentration {c}')
        flux, total_flux = solve_problem_instance(float(c), t_final, dt, mesh, bdry, sigma, gamma, [], V1, D1, D2,
                                                  None, steady_state)
        finite_ele_fluxes.append(total_flux)
    finite_ele_fluxes = np.array(finite_ele_fluxes)
    print(f'Finite fluxes: {finite_ele_fluxes} required rates: {required_rate(c_vals, sigma_orig, gamma_orig)} '
          f'residuals: {finite_ele_fluxes - required_rate(c_vals, sigma_orig, gamma_orig)}')
    return finite_ele_fluxes - required_rate(c_vals, sigma_orig, gamma_orig)


def main(n=32, max_nfev=6, c_vals=(1, 2, 4, 8)):
    sigma_orig, gamma_orig = 0.1, 1
    r1_max = r2_max = 5
    D_A = D_B = 1
    D_C = 5
    D1 = D_A + D_B
    D2 = D_C + 1 / (1 / D_A + 1 / D_B)
    dt, t_final, steady_state = 0.1, 0.1, True
    c_vals = np.array(c_vals, dtype=float)
    init_guess = np.array([0.09867282, 1.20166547])
    bounds = ([1e-10, 1e-10], [np.inf, np.inf])
    start_time = time.time()
    sol = least_squares(fun=finite_element_fluxes,
                        args=[sigma_orig, gamma_orig, c_vals, r1_max, r2_max, n, t_final, dt, D1, D2, steady_state],
                        x0=init_guess, jac=jacobian, method='trf', bounds=bounds, max_nfev=max_nfev,
                        x_scale=[sigma_orig / 10, gamma_orig / 10])
    print('******************************************************************************')
    print(f'Elapsed time: {time.time() - start_time}')
    print(f'Solution parameters: {sol.x} cost: {sol.cost} residuals: {sol.fun}')
    return sol


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 6)
