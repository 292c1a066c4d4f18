class Optimizer: pass
