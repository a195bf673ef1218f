// MirageSphericalSigmoid (src/functions.rs:67-82) restated with 4 special-function evaluations per call instead of 6:
//   * normalize + length share one inverseSqrt
//   * sigmoid(x) = 1/(1+E), E = exp(-x): t = 200 - 185/(1+E); n(t) = A(t)/B(t) with A, B linear in t, so in
//     eta = n(t_in)/n(t_out) every 1/(1+E) cancels: eta = A'(E_in) B'(E_out) / (A'(E_out) B'(E_in)), A', B' linear in E
//   * the clamp(0, 1, sigmoid) = min(1, sigmoid) of the reference never binds (sigmoid <= 1)
let center = vec3<f32>(-0.5, 0.5, -0.5);
let c1p = 0.0000104 * 101325.0;
let a0 = c1p * (1.0 + 101325.0 * 0.0000000001 * 60.1);
let a1 = c1p * 101325.0 * 0.0000000001 * 0.972;
let c2 = 0.00366;
let A1 = a0 - 15.0 * a1;
let A2 = a0 - 200.0 * a1;
let B1 = 1.0 + 15.0 * c2;
let B2 = 1.0 + 200.0 * c2;
let ks = -48.0 * 1.4426950408889634;
let k0 = 6.0 * 1.4426950408889634;
let cd = p - center;
let d2 = dot(cd, cd);
let r = inverseSqrt(d2);
let normal = cd * r;
let cdi = p_prev - center;
let d2i = dot(cdi, cdi);
let e_in = exp2(d2i * inverseSqrt(d2i) * ks + k0);
let e_out = exp2(d2 * r * ks + k0);
let eta = ((A1 + A2 * e_in) * (B1 + B2 * e_out)) / ((A1 + A2 * e_out) * (B1 + B2 * e_in));
let d = dot(v, normal);
let cr = min(1.0, d);
let k = 1.0 - eta * eta * (1.0 - cr * cr);
let sk = sqrt(k);
let cs = select(sk, -sk, cr < 0.0) - eta * cr;
let refr = eta * v + cs * normal;
return select(refr, v - (2.0 * d) * normal, k < 0.0);
