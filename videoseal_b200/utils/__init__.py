"""Mirrors the reference's `videoseal.utils` import path for the loader entry points (`from videoseal.utils.cfg import ...`)."""
