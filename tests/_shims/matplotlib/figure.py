class Figure: pass
