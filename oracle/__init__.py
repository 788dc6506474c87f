"""CPU oracle for the hot path -- TEST INFRASTRUCTURE, never imported by the product package."""
