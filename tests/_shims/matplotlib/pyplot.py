def gcf(): pass
