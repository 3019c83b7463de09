"""Rayleigh-Benard residual layer.  Drop-in for reference experiments/rb2d/physics.py:6-64.

Variables: inputs (t, x, z) bind to coordinate columns 0, 1, 2 *by position*; outputs (p, b, u, w).
With P = (Ra Pr)^-1/2, R = (Ra / Pr)^-1/2 and n_* = 1 / crop_*:

    b :  n_t b_t - P (n_x^2 b_xx + n_z^2 b_zz)        + (u n_x b_x + w n_z b_z)
    u :  n_t u_t - R (n_x^2 u_xx + n_z^2 u_zz) + p_x     + (u n_x u_x + w n_z u_z)
    w :  n_t w_t - R (n_x^2 w_xx + n_z^2 w_zz) + p_z - b + (u n_x w_x + w n_z w_z)
    continuity (optional):  n_x u_x + n_z w_z
"""
from .pde import PDELayer


def get_rb2_pde_layer(mean=None, std=None, t_crop=2., z_crop=1., x_crop=2., prandtl=1., rayleigh=1e6,
                      use_continuity=False):
    """Build the PDELayer of the RB2 governing equations (forward method still to be set)."""
    P = (rayleigh * prandtl)**(-1/2)
    R = (rayleigh / prandtl)**(-1/2)
    in_vars = 't, x, z'
    out_vars = 'p, b, u, w'
    nt, nz, nx = 1./t_crop, 1./z_crop, 1./x_crop

    def transport(var, nu, source):
        diffusion = f'{nu}*(({nx})**2*dif(dif({var},x),x)+({nz})**2*dif(dif({var},z),z))'
        advection = f'(u*{nx}*dif({var},x)+w*{nz}*dif({var},z))'
        return f'{nt}*dif({var},t)-{diffusion}{source}+{advection}'

    equations = [('transport_eqn_b', transport('b', P, '')),
                 ('transport_eqn_u', transport('u', R, '+dif(p,x)')),
                 ('transport_eqn_w', transport('w', R, '+dif(p,z)-b'))]
    if use_continuity:
        equations.append(('continuity', f'{nx} * dif(u, x) + {nz} * dif(w, z)'))

    subs_dict = None
    if (mean is not None) or (std is not None):
        if not ((mean is not None) and (std is not None)):
            raise ValueError('mean and std must either be both None, or both arrays of len 4.')
        if not (hasattr(mean, '__len__') and hasattr(std, '__len__')):
            raise TypeError("mean and std must be arrays of len 4. instead they are {} and {}"
                            .format(type(mean), type(std)))
        if not (len(mean) == 4 and len(std) == 4):
            raise ValueError("mean and std must be arrays of len 4. instead they are of len {} and {}"
                             .format(len(mean), len(std)))
        names = [v.strip() for v in out_vars.split(',')]
        subs_dict = {v: f"{v}*{std[i]}+{mean[i]}" for i, v in enumerate(names)}

    pde_layer = PDELayer(in_vars=in_vars, out_vars=out_vars)
    for name, eqn in equations:
        pde_layer.add_equation(eqn, name, subs_dict=subs_dict)
    return pde_layer
