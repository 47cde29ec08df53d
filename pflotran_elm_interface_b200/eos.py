"""Water density for decks that do not set REFERENCE_LIQUID_DENSITY.

Transport-only runs take ``global_auxvar%den_kg`` from
``EOSWaterDensity(reference_temperature, reference_pressure)`` with the default
IFC67 formulation (src/pflotran/eos.F90:706-711, option_flow.F90:137-141,
eos_water.F90:1583-1708).  Host set-up only.
"""

H2O_CRITICAL_TEMPERATURE = 647.3   # K, pflotran_constants.F90:76
H2O_CRITICAL_PRESSURE = 22.064e6   # Pa, pflotran_constants.F90:80

_AA = [6.824687741e03, -5.422063673e02, -2.096666205e04, 3.941286787e04, -6.733277739e04, 9.902381028e04,
       -1.093911774e05, 8.590841667e04, -4.511168742e04, 1.418138926e04, -2.017271113e03, 7.982692717e00,
       -2.616571843e-2, 1.522411790e-3, 2.284279054e-2, 2.421647003e02, 1.269716088e-10, 2.074838328e-7,
       2.174020350e-8, 1.105710498e-9, 1.293441934e01, 1.308119072e-5, 6.047626338e-14]
_A = [None, 8.438375405e-1, 5.362162162e-4, 1.720000000e00, 7.342278489e-2, 4.975858870e-2, 6.537154300e-1,
      1.150000000e-6, 1.510800000e-5, 1.418800000e-1, 7.002753165e00, 2.995284926e-4, 2.040000000e-1]


def water_density_ifc67(t_c: float = 25.0, p_pa: float = 101325.0) -> float:
    """kg/m^3 (eos_water.F90:1662-1707)"""
    aa, a = _AA, _A
    vc1 = 0.00317
    theta = (t_c + 273.15) / H2O_CRITICAL_TEMPERATURE
    theta2x = theta * theta
    theta18 = theta ** 18.0
    theta20 = theta18 * theta2x
    beta = p_pa / H2O_CRITICAL_PRESSURE
    beta2x = beta * beta
    yy = 1.0 - a[1] * theta2x - a[2] * theta ** (-6.0)
    xx = a[3] * yy * yy - 2.0 * (a[4] * theta - a[5] * beta)
    xx = xx ** 0.5 if xx > 0.0 else 1.0e-6
    zz = yy + xx
    u0 = -5.0 / 17.0
    u1 = aa[11] * a[5] * zz ** u0
    u2 = 1.0 / (a[8] + theta ** 11.0)
    u3 = aa[17] + (2.0 * aa[18] + 3.0 * aa[19] * beta) * beta
    u4 = 1.0 / (a[7] + theta18 * theta)
    u5 = (a[10] + beta) ** (-4.0)
    u6 = a[11] - 3.0 * u5
    u7 = aa[20] * theta18 * (a[9] + theta2x)
    u8 = aa[15] * (a[6] - theta) ** 9.0
    vr = (u1 + aa[12] + theta * (aa[13] + aa[14] * theta) + u8 * (a[6] - theta) + aa[16] * u4 - u2 * u3 - u6 * u7
          + (3.0 * aa[21] * (a[12] - theta) + 4.0 * aa[22] * beta / theta20) * beta2x)
    return 1.0 / (vr * vc1)
