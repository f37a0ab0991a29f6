import sympy as sp
u, v = sp.symbols("u v")
a, b, c, p, q, r = sp.symbols("a b c p q r")   # a=d12^2, b=d13^2, c=d23^2; p=cos12, q=cos13, r=cos23
w = 1 + u**2 - 2*u*p
A = b*w - a*(1 + v**2 - 2*v*q)
B = c*w - a*(u**2 + v**2 - 2*u*v*r)
# A - B is linear in v:
lin = sp.expand(A - B)
v_num = -lin.coeff(v, 0); v_den = lin.coeff(v, 1)
print("v = (", sp.simplify(v_num), ") / (", sp.simplify(v_den), ")")
# substitute into A * den^2
quart = sp.expand(sp.together(A.subs(v, v_num / v_den) * v_den**2))
poly = sp.Poly(quart, u)
print("degree", poly.degree())
cs = poly.all_coeffs()
for i, cf in enumerate(cs):
    print(f"k{poly.degree()-i} =", sp.factor(cf))
import pickle
pickle.dump([str(sp.factor(cf)) for cf in cs], open("p3p_coeffs.pkl", "wb"))
