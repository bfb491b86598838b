"""Host-side geometry value types with gtsam's names (Rot3, Pose3, Point3, StereoPoint2, Cal3_S2Stereo).

These are light carriers used while building the graph on the host, exactly where
/root/reference/batch.py uses the gtsam classes (batch.py:46-47, :83, :115, :131-135, :190,
:300-301).  No optimisation arithmetic runs here: residuals, Jacobians and retraction live in
the CUDA kernels under csrc/.  Tangent convention follows gtsam: xi = [omega; v].
"""
import math
import numpy as np


def Point3(x=0.0, y=0.0, z=0.0):
    """gtsam >= 4.1 returns Point3 as a plain 3-vector (batch.py:46, :132)."""
    return np.array([x, y, z], dtype=np.float64)


def _hat(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


class Rot3:
    __slots__ = ("_R",)

    def __init__(self, R=None):
        self._R = np.eye(3) if R is None else np.array(R, dtype=np.float64).reshape(3, 3)

    # --- constructors used by the reference
    @staticmethod
    def Quaternion(w, x, y, z):
        """w first (batch.py:47, :131)."""
        nrm = math.sqrt(w * w + x * x + y * y + z * z)
        w, x, y, z = w / nrm, x / nrm, y / nrm, z / nrm
        return Rot3([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    @staticmethod
    def Rodrigues(wx, wy=None, wz=None):
        """Rot3.Rodrigues(wx, wy, wz) or Rot3.Rodrigues(vec3) (batch.py:190)."""
        w = np.array([wx, wy, wz], dtype=np.float64) if wy is not None else np.asarray(wx, dtype=np.float64)
        return Rot3.Expmap(w)

    @staticmethod
    def Expmap(w):
        w = np.asarray(w, dtype=np.float64).reshape(3)
        th2 = float(w @ w)
        W = _hat(w)
        if th2 <= np.finfo(float).eps:
            return Rot3(np.eye(3) + W)
        th = math.sqrt(th2)
        K = W / th
        return Rot3(np.eye(3) + math.sin(th) * K + (2.0 * math.sin(0.5 * th) ** 2) * (K @ K))

    @staticmethod
    def Ypr(y, p, r):
        cy, sy, cp, sp, cr, sr = math.cos(y), math.sin(y), math.cos(p), math.sin(p), math.cos(r), math.sin(r)
        Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
        Ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
        Rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
        return Rot3(Rz @ Ry @ Rx)

    def matrix(self):
        return self._R.copy()

    def transpose(self):
        return self._R.T.copy()

    def inverse(self):
        return Rot3(self._R.T)

    def compose(self, other):
        return Rot3(self._R @ other._R)

    __mul__ = compose

    def rotate(self, p):
        return self._R @ np.asarray(p, dtype=np.float64)

    def unrotate(self, p):
        return self._R.T @ np.asarray(p, dtype=np.float64)

    def __repr__(self):
        return f"Rot3(\n{self._R}\n)"


class Pose3:
    __slots__ = ("_R", "_t")

    def __init__(self, R=None, t=None):
        if isinstance(R, Pose3):
            self._R, self._t = R._R.copy(), R._t.copy()
            return
        if R is not None and t is None and not isinstance(R, Rot3):
            M = np.asarray(R, dtype=np.float64)
            if M.shape == (4, 4):
                self._R, self._t = M[:3, :3].copy(), M[:3, 3].copy()
                return
        self._R = R._R.copy() if isinstance(R, Rot3) else (np.eye(3) if R is None else np.array(R, float).reshape(3, 3))
        self._t = np.zeros(3) if t is None else np.array(t, dtype=np.float64).reshape(3)

    def rotation(self):
        return Rot3(self._R)

    def translation(self):
        return self._t.copy()

    def x(self):
        return float(self._t[0])

    def y(self):
        return float(self._t[1])

    def z(self):
        return float(self._t[2])

    def matrix(self):
        M = np.eye(4)
        M[:3, :3] = self._R
        M[:3, 3] = self._t
        return M

    def compose(self, o):
        return Pose3(Rot3(self._R @ o._R), self._t + self._R @ o._t)

    __mul__ = compose

    def inverse(self):
        return Pose3(Rot3(self._R.T), -self._R.T @ self._t)

    def between(self, o):
        return self.inverse().compose(o)

    def transformFrom(self, p):
        return self._R @ np.asarray(p, float) + self._t

    def transformTo(self, p):
        return self._R.T @ (np.asarray(p, float) - self._t)

    @staticmethod
    def Expmap(xi):
        xi = np.asarray(xi, dtype=np.float64).reshape(6)
        w, v = xi[:3], xi[3:]
        R = Rot3.Expmap(w)
        th2 = float(w @ w)
        if th2 > np.finfo(float).eps:
            wxv = np.cross(w, v)
            t = (wxv - R._R @ wxv + w * float(w @ v)) / th2
        else:
            t = v.copy()
        return Pose3(R, t)

    def retract(self, xi):
        return self.compose(Pose3.Expmap(xi))

    def as_row(self):
        """12 doubles: R row-major then t (the host AoS row the packer transposes to SoA)."""
        return np.concatenate([self._R.reshape(9), self._t])

    @staticmethod
    def from_row(row):
        row = np.asarray(row, dtype=np.float64)
        return Pose3(Rot3(row[:9].reshape(3, 3)), row[9:12])

    def equals(self, o, tol=1e-9):
        return bool(np.allclose(self._R, o._R, atol=tol) and np.allclose(self._t, o._t, atol=tol))

    def __repr__(self):
        return f"Pose3(R=\n{self._R},\n t={self._t})"


class StereoPoint2:
    __slots__ = ("_v",)

    def __init__(self, uL=0.0, uR=0.0, v=0.0):
        self._v = np.array([uL, uR, v], dtype=np.float64)

    def uL(self):
        return float(self._v[0])

    def uR(self):
        return float(self._v[1])

    def v(self):
        return float(self._v[2])

    def vector(self):
        return self._v.copy()


class Cal3_S2Stereo:
    """Cal3_S2Stereo(fx, fy, s, u0, v0, b) (batch.py:115)."""
    __slots__ = ("_k",)

    def __init__(self, fx=1.0, fy=1.0, s=0.0, u0=0.0, v0=0.0, b=1.0):
        self._k = np.array([fx, fy, s, u0, v0, b], dtype=np.float64)

    def fx(self):
        return float(self._k[0])

    def fy(self):
        return float(self._k[1])

    def skew(self):
        return float(self._k[2])

    def px(self):
        return float(self._k[3])

    def py(self):
        return float(self._k[4])

    def baseline(self):
        return float(self._k[5])

    def vector(self):
        return self._k.copy()
