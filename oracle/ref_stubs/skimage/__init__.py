"""Restated slice of scikit-image 0.18.1 (Dockerfile:69 pin) — only skimage.draw.disk is on the path."""
