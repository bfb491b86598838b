class LevenbergMarquardtParams: pass
class LevenbergMarquardtOptimizer: pass
