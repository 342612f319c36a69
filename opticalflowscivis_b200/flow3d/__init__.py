"""Drop-in for the reference's Flow-3D package (model.warplayer / model.IFNet / model.RIFE)."""
