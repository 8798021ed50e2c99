"""Host-side mirror of the reference's ``src/utils`` modules that sit on the LEC hot path."""
