"""Drop-in for the reference's Flow-2D package (model.warplayer / model.IFNet / model.RIFE)."""
