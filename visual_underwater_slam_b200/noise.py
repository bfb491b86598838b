"""gtsam.noiseModel.{Diagonal,Isotropic,Gaussian} carriers (batch.py:95-98, :118, :189).

Whitening itself (A <- Sigma^-1/2 A, b <- Sigma^-1/2 b; gtsam/linear/NoiseModel.cpp) is fused
into the CUDA linearize kernels; these objects only hold the square-root information that the
packer copies into the per-factor tables.
"""
import numpy as np


class _Base:
    def dim(self):
        return self._dim

    def sigmas(self):
        raise NotImplementedError

    def sqrt_info_diag(self):
        """1/sigma per row, or None if the model is not diagonal."""
        return None

    def R(self):
        """Upper-triangular R with R^T R = Sigma^-1 (gtsam Gaussian::R())."""
        return np.diag(self.sqrt_info_diag())


class _Diagonal(_Base):
    def __init__(self, sigmas):
        self._sigmas = np.asarray(sigmas, dtype=np.float64).reshape(-1).copy()
        self._dim = self._sigmas.size
        if np.any(self._sigmas <= 0):
            raise ValueError("constrained (sigma<=0) noise models are not supported on this path")

    def sigmas(self):
        return self._sigmas.copy()

    def sqrt_info_diag(self):
        return 1.0 / self._sigmas


class _Gaussian(_Base):
    def __init__(self, R):
        self._R = np.asarray(R, dtype=np.float64)
        self._dim = self._R.shape[0]

    def R(self):
        return self._R.copy()

    def covariance(self):
        Ri = np.linalg.inv(self._R)
        return Ri @ Ri.T

    def sigmas(self):
        return np.sqrt(np.diag(self.covariance()))


class Diagonal:
    @staticmethod
    def Sigmas(sigmas):
        return _Diagonal(sigmas)

    @staticmethod
    def Variances(v):
        return _Diagonal(np.sqrt(np.asarray(v, dtype=np.float64)))

    @staticmethod
    def Precisions(p):
        return _Diagonal(1.0 / np.sqrt(np.asarray(p, dtype=np.float64)))


class Isotropic:
    @staticmethod
    def Sigma(dim, sigma):
        return _Diagonal(np.full(int(dim), float(sigma)))

    @staticmethod
    def Variance(dim, variance):
        return _Diagonal(np.full(int(dim), float(np.sqrt(variance))))


class Unit:
    @staticmethod
    def Create(dim):
        return _Diagonal(np.ones(int(dim)))


def upper_sqrt_information(cov):
    """R upper, R^T R = cov^-1, batched [n,d,d] -> [n,d,d] (Gaussian::Covariance, NoiseModel.cpp)."""
    cov = np.asarray(cov, dtype=np.float64)
    info = np.linalg.inv(cov)
    info = 0.5 * (info + np.swapaxes(info, -1, -2))
    return np.swapaxes(np.linalg.cholesky(info), -1, -2)


class Gaussian:
    @staticmethod
    def Covariance(cov):
        return _Gaussian(upper_sqrt_information(np.asarray(cov, dtype=np.float64)))

    @staticmethod
    def Information(info):
        info = np.asarray(info, dtype=np.float64)
        return _Gaussian(np.linalg.cholesky(0.5 * (info + info.T)).T)

    @staticmethod
    def SqrtInformation(R):
        return _Gaussian(R)


class noiseModel:  # namespace object: gtsam.noiseModel.Diagonal.Sigmas(...)
    Diagonal = Diagonal
    Isotropic = Isotropic
    Gaussian = Gaussian
    Unit = Unit
