-- depth_estimation_api_patch.lua -- what changes in depth_estimation_api.lua (reference lines
-- 164-170) to run the matching path on the GPU.  Everything else in nextFrameDepth() stays.
-- NOT EXECUTED in this repository (no Lua in the build image).
--
--   before (depth_estimation_api.lua:164-170):
--      local input   = prepareInput(geometry, last_filtered, filtered)
--      local moutput = model:forward(input)
--      local poutput = processOutput(geometry, moutput, true, nil)
--      output = poutput.full
--
--   after:
require 'nn_depthmatch'
local dense = nn.DenseMatch(geometry)                 -- once, next to loadModel()

local function match(geometry, last_filtered, filtered)
   local input = prepareInput(geometry, last_filtered, filtered)   -- still a narrow()ed view
   local r = dense:forward(input)                     -- one dm_match_extract call
   local yoff, xoff = centered2onebased(geometry, 0, 0)
   local poutput = {index = r.index, full = r.full,
                    y = r.soft[1] - yoff, x = r.soft[2] - xoff,   -- 'mean' extraction
                    full_confidences = nil}
   return poutput
end
return match
