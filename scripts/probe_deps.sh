#!/bin/bash
# Probe the GPU box for the reference's un-vendored third-party stack.
# Output is committed under profiles/ as the transcript of what a real-Box2D pin could (not) use.
set +e
echo "== date"; date -u
echo "== python"; which python; python --version
for m in Box2D gym gymnasium pyglet shapely pygame OpenGL pymunk gym_multi_car_racing; do
  python - <<PY
import importlib
try:
    m = importlib.import_module("$m"); print("import $m: OK", getattr(m, "__version__", "?"), getattr(m, "__file__", "?"))
except Exception as e:
    print("import $m: FAIL", type(e).__name__, e)
PY
done
echo "== pip download (no network expected)"
timeout 60 python -m pip download --no-deps -d /tmp/pipdl box2d-py==2.3.5 2>&1 | tail -3
timeout 60 python -m pip download --no-deps -d /tmp/pipdl gym==0.17.2 2>&1 | tail -3
echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "box2d|gym|pyglet|shapely|pygame|opengl" ; echo "(end wheelhouse grep)"
echo "== baseline/_ref"; ls -la baseline/_ref 2>&1 | head
echo "== pip list grep"; python -m pip list 2>/dev/null | grep -i -E "box2d|gym|pyglet|shapely|pygame|opengl|mujoco|pymunk"; echo "(end pip grep)"
echo "== find"; timeout 120 find / -xdev \( -iname "*box2d*" -o -iname "car_dynamics*" -o -iname "car_racing*" -o -iname "b2World*" -o -iname "libGL.so*" -o -iname "libEGL.so*" -o -iname "libOSMesa*" -o -iname "Xvfb" -o -iname "libgeos*" \) 2>/dev/null | grep -v "^/proc" | head -40; echo "(end find)"
echo "== swig/X"; which swig Xvfb xvfb-run glxinfo 2>&1
